"""Pin the oracle: backbone restatement vs golden vectors made from the reference's own
backbone.py (tests/golden/make_golden.py); GP restatement vs scikit-learn (independent
implementation) and vs closed-form gradients."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import backbone as obb
from oracle import episode as oep
from oracle import gp as ogp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _checksum(p):
    return sum(float(p[k].double().abs().sum()) for k in sorted(p) if p[k].is_floating_point())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "backbone_*.npz"))))
def test_backbone_matches_reference_golden(path):
    name = os.path.basename(path)[len("backbone_"):-len(".npz")]
    bn_out = name.endswith("_bnout")
    arch = name.replace("_bnout", "")
    gold = np.load(path)
    p = obb.init_params(arch, seed=3, bn_out=bn_out)
    assert abs(_checksum(p) - float(gold["checksum"])) < 1e-6 * float(gold["checksum"]), "RNG drift"
    g = torch.Generator().manual_seed(7)
    x = torch.randn(int(gold["n"]), 3, int(gold["size"]), int(gold["size"]), generator=g)
    out_train = obb.forward(arch, p, x, training=True)
    np.testing.assert_allclose(out_train.numpy(), gold["out_train"], rtol=1e-4, atol=1e-5)
    for k in gold.files:
        if k.startswith("stat_"):
            np.testing.assert_allclose(p[k[5:]].numpy(), gold[k], rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        out_eval = obb.forward(arch, p, x, training=False)
    np.testing.assert_allclose(out_eval.numpy(), gold["out_eval"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("kernel", ["linear", "rbf"])
def test_gp_matches_sklearn(kernel):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, WhiteKernel
    rng = np.random.RandomState(0)
    n, m, d = 25, 40, 16
    x = rng.randn(n, d)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    xs = rng.randn(m, d)
    xs /= np.linalg.norm(xs, axis=1, keepdims=True)
    y = np.where(np.arange(n) < 5, 1.0, -1.0)
    p = ogp.default_gp_params(kernel, 1, d, dtype=torch.float64)
    s = float(torch.nn.functional.softplus(p["raw_outputscale"][0]))
    sig2 = float(ogp.noise(p, 0))
    if kernel == "linear":
        v = float(torch.nn.functional.softplus(p["raw_variance"][0]))
        k = ConstantKernel(s * v, "fixed") * DotProduct(sigma_0=0.0, sigma_0_bounds="fixed")
    else:
        ls = float(torch.nn.functional.softplus(p["raw_lengthscale"][0]))
        k = ConstantKernel(s, "fixed") * RBF(ls, "fixed")
    k = k + WhiteKernel(sig2, "fixed")
    gpr = GaussianProcessRegressor(kernel=k, optimizer=None, alpha=0.0).fit(x, y)
    tx, txs, ty = torch.from_numpy(x), torch.from_numpy(xs), torch.from_numpy(y)
    lm = float(ogp.log_marginal(kernel, tx, ty, p, 0))
    assert abs(lm - gpr.log_marginal_likelihood_value_) < 1e-9 * max(1.0, abs(lm))
    mean, var = ogp.predict(kernel, tx, ty[None], txs, p, want_var=True)
    mu, std = gpr.predict(xs, return_std=True)
    np.testing.assert_allclose(mean[0].numpy(), mu, rtol=1e-9, atol=1e-11)
    # sklearn's predictive std includes the WhiteKernel noise, like likelihood(model(x))
    np.testing.assert_allclose(var[0].sqrt().numpy(), std, rtol=1e-8, atol=1e-10)


def test_mll_gradient_closed_form():
    """autograd through Cholesky == (K^-1 - alpha alpha^T)/(2NC) etc. (Appendix A)."""
    torch.manual_seed(1)
    n, d, c = 12, 7, 3
    z = torch.nn.functional.normalize(torch.randn(n, d, dtype=torch.float64), dim=1).requires_grad_(True)
    p = ogp.default_gp_params("bncossim", c, d, dtype=torch.float64)
    p["raw_outputscale"] = torch.tensor([0.3, -0.2, 0.8], dtype=torch.float64, requires_grad=True)
    p["constant"] = torch.tensor([0.1, -0.3, 0.0], dtype=torch.float64, requires_grad=True)
    t = oep.make_targets(c, n // c, torch.float64)
    loss = ogp.mll_loss("bncossim", z, t, p)
    loss.backward()
    g = (z @ z.t()).detach()
    dz = torch.zeros_like(g)
    for ci in range(c):
        s = torch.nn.functional.softplus(p["raw_outputscale"][ci]).detach()
        kt = s * g + 0.1 * torch.eye(n, dtype=torch.float64)
        kinv = torch.linalg.inv(kt)
        r = t[ci] - p["constant"][ci].detach()
        alpha = kinv @ r
        dk = (kinv - torch.outer(alpha, alpha)) / (2 * n * c)
        ds = (dk * g).sum() * torch.sigmoid(p["raw_outputscale"][ci].detach())
        assert abs(ds - p["raw_outputscale"].grad[ci]) < 1e-10
        assert abs(-(alpha.sum()) / (n * c) - p["constant"].grad[ci]) < 1e-10
        dz += s * dk
    np.testing.assert_allclose(((dz + dz.t()) @ z.detach()).numpy(), z.grad.numpy(), rtol=1e-8, atol=1e-12)
    expect = math.log(2 * math.pi)  # sanity: loss contains the N log 2pi / (2N) term
    assert loss.item() > 0.5 * expect - 5


def test_oracle_train_step_decreases_loss():
    torch.manual_seed(0)
    o = oep.OracleDKT("Conv4", "bncossim", n_way=3, n_support=2, seed=0)
    x = oep.synthetic_episode(0, n_way=3, n_support=2, n_query=3, image_size=84)
    first = o.train_step(x)
    for _ in range(3):
        last = o.train_step(x)
    assert float(last["loss"][0]) < float(first["loss"][0])
    assert first["mean_query"].shape == (1, 3, 9)
    ok, cnt, _ = o.correct(x)
    assert cnt == 9 and 0 <= ok <= 9


def test_oracle_regression_step():
    torch.manual_seed(0)
    for kernel in ("rbf", "spectral"):
        o = oep.OracleDKTRegression(kernel)
        x = torch.randn(6, 3, 100, 100)
        y = torch.linspace(-1, 1, 6)
        r = o.train_step(x, y)
        assert torch.isfinite(r["loss"])
        mse, mean, lo, hi = o.test_episode(x[:3], y[:3], x, y)
        assert mean.shape == (6,) and bool((hi > lo).all())


# ----------------------------------------------------------------------------- control flow pinned to the reference
# tests/golden/dkt_*.npz were produced by running the reference's own, unmodified methods/DKT.py and
# methods/DKT_regression.py (tests/golden/make_golden_dkt.py, on oracle/gpytorch_standin because GPyTorch cannot be
# installed offline).  The stand-in's arithmetic is independent of oracle/gp.py (scipy / LAPACK in float64 + analytic
# gradients, validated against scikit-learn when the fixtures are made), so these tests pin BOTH oracle/episode.py's
# restatement of train_loop / correct / test_loop / get_logits AND oracle/gp.py's arithmetic against a second evaluation.
# Quantities behind Adam steps: the conv biases cancel inside BatchNorm, their gradients are rounding noise and Adam
# moves them by ~+-lr per step whatever the sign -- trunk.*.C.bias and what depends on it (the first running mean, the
# eval-mode logits of the trained model) are therefore compared at the size of those moves, everything else tightly.
def test_standin_does_not_use_the_oracle():
    src = open(os.path.join(os.path.dirname(GOLD), "..", "oracle", "gpytorch_standin", "gpytorch", "__init__.py")).read()
    assert "import gp" not in src and "from oracle" not in src and "ogp." not in src



def _golden_episodes(count, seed, n_way=3, per_class=5, image=84):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        x = torch.randn(n_way, per_class, 3, image, image, generator=g)
        out.append(x + 0.5 * torch.randn(n_way, 1, 3, 1, 1, generator=g))
    return out


def _golden_gp_start(kernel, n_models):
    out = {"constant": torch.tensor([0.05 * (c + 1) for c in range(n_models)]),
           "raw_outputscale": torch.tensor([0.1 * c - 0.1 for c in range(n_models)])}
    if kernel == "rbf":
        out["raw_lengthscale"] = torch.tensor([30.0 + 5.0 * c for c in range(n_models)])
    if kernel == "linear":
        out["raw_variance"] = torch.tensor([-3.0 + 0.3 * c for c in range(n_models)])
    return out


@pytest.mark.parametrize("kernel", ["bncossim", "rbf", "cossim", "linear"])
def test_episode_oracle_matches_reference_control_flow(kernel):
    gold = np.load(os.path.join(GOLD, "dkt_cls_%s.npz" % kernel))
    torch.set_num_threads(1)
    o = oep.OracleDKT("Conv4", kernel, n_way=3, n_support=2, seed=5)
    for k, v in _golden_gp_start(kernel, 3).items():
        o.gp[k] = v.clone()
    tol = dict(rtol=2e-5, atol=2e-6)
    test_eps = _golden_episodes(2, seed=13, per_class=6)
    np.testing.assert_allclose(np.stack([o.get_logits(x).numpy() for x in test_eps]), gold["logits_init"], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(np.array([o.correct(x) for x in test_eps], dtype=np.float64), gold["correct_init"])
    res = [o.train_step(x) for x in _golden_episodes(3, seed=11)]
    np.testing.assert_allclose([float(r["loss"][0]) for r in res], gold["loss"], **tol)
    np.testing.assert_array_equal([r["acc_support"][0] for r in res], gold["acc_support"])
    np.testing.assert_array_equal([r["acc_query"][0] for r in res], gold["acc_query"])
    # three Adam steps later: same weights, running statistics and hyper-parameters
    ptol = dict(rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(o.bb["trunk.0.C.bias"].detach().numpy(), gold["after_conv0_bias"], atol=3.5e-3)      # 3 steps of lr 1e-3
    np.testing.assert_allclose(o.bb["trunk.3.C.weight"].detach().numpy().reshape(-1)[:256], gold["after_conv3_weight_head"], **ptol)
    np.testing.assert_allclose(o.bb["trunk.0.BN.running_mean"].numpy(), gold["after_bn0_running_mean"], atol=1.1e-3)   # momentum 0.1 x bias moves
    np.testing.assert_allclose(o.bb["trunk.3.BN.running_var"].numpy(), gold["after_bn3_running_var"], **ptol)
    if kernel == "bncossim":
        np.testing.assert_allclose(o.bb["trunk.bn_out.weight"].detach().numpy()[:256], gold["after_bn_out_weight_head"], **ptol)
        np.testing.assert_allclose(o.bb["trunk.bn_out.running_mean"].numpy()[:256], gold["after_bn_out_running_mean_head"], **ptol)
    total = sum(float(v.double().abs().sum()) for v in o.bb.values() if v.is_floating_point())
    assert abs(total - float(gold["after_param_abs_sum"])) < 2e-4 * total
    for k in gold.files:
        if k.startswith("after_gp_"):
            np.testing.assert_allclose(o.gp[k[len("after_gp_"):]].detach().numpy(), gold[k], rtol=1e-5, atol=1e-7)
    # a new train_loop call: Adam restarts (DKT.py:114-115)
    o.new_optimizer()
    r = o.train_step(_golden_episodes(1, seed=12)[0])
    np.testing.assert_allclose([float(r["loss"][0])], gold["loss_second_call"], **tol)
    # test path: get_logits, test_loop (mean / std of per-episode accuracy), correct, correct(N=3)
    logits = np.stack([o.get_logits(x).numpy() for x in test_eps])
    assert np.abs(logits - gold["logits"]).max() <= 1e-2 * np.abs(gold["logits"]).max()      # behind the conv-bias moves
    corr = np.array([o.correct(x) for x in test_eps], dtype=np.float64)
    np.testing.assert_array_equal(corr, gold["correct"])
    acc = corr[:, 0] / corr[:, 1] * 100
    assert abs(acc.mean() - float(gold["test_acc_mean"])) < 1e-9 and abs(acc.std() - float(gold["test_acc_std"])) < 1e-9
    adapt = np.array(o.correct(test_eps[0], N=3), dtype=np.float64)
    np.testing.assert_array_equal(adapt[:2], gold["correct_adapt"][:2])
    np.testing.assert_allclose(adapt[2], gold["correct_adapt"][2], rtol=1e-2)      # loss of the trained model: behind the conv-bias moves
    for k in gold.files:
        if k.startswith("adapt_gp_"):      # 3 more Adam steps on the trained model's features
            np.testing.assert_allclose(o.gp[k[len("adapt_gp_"):]].detach().numpy(), gold[k], rtol=1e-4, atol=5e-6)
    la = o.get_logits(test_eps[1]).numpy()
    assert np.abs(la - gold["logits_after_adapt"]).max() <= 1e-2 * np.abs(gold["logits_after_adapt"]).max()


@pytest.mark.parametrize("kernel", ["rbf", "spectral"])
def test_regression_oracle_matches_reference_control_flow(kernel):
    gold = np.load(os.path.join(GOLD, "dkt_reg_%s.npz" % kernel))
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(21)
    batch = torch.rand(3, 19, 3, 100, 100, generator=g)
    labels = torch.rand(3, 19, generator=g) * 2 - 1
    o = oep.OracleDKTRegression(kernel, seed=5)
    losses = [float(o.train_step(x, y)["loss"]) for x, y in zip(batch, labels)]      # DKT_regression.py:48-57
    np.testing.assert_allclose(losses, gold["loss"], rtol=2e-5)
    ptol = dict(rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(o.bb["layer1.bias"].detach().numpy(), gold["after_layer1_bias"], **ptol)
    np.testing.assert_allclose(o.gp["constant"].detach().numpy(), gold["after_constant"], **ptol)
    np.testing.assert_allclose(o.gp["raw_noise"].detach().numpy(), gold["after_raw_noise"], **ptol)
    if kernel == "rbf":
        np.testing.assert_allclose(o.gp["raw_outputscale"].detach().numpy(), gold["after_raw_outputscale"], **ptol)
        np.testing.assert_allclose(o.gp["raw_lengthscale"].detach().numpy(), gold["after_raw_lengthscale"], **ptol)
    else:
        np.testing.assert_allclose(o.gp["raw_mixture_weights"].detach().numpy(), gold["after_raw_mixture_weights"], **ptol)
        np.testing.assert_allclose(o.gp["raw_mixture_means"].detach().numpy().reshape(4, -1)[:, :64],
                                   gold["after_raw_mixture_means_head"], **ptol)
    ind, n = [int(i) for i in gold["test_support_ind"]], int(gold["test_person"])
    mse, mean, lo, hi = o.test_episode(batch[n][ind], labels[n][ind], batch[n], labels[n])   # DKT_regression.py:66-95
    np.testing.assert_allclose(float(mse), float(gold["test_mse"]), rtol=1e-4)


def test_sines_oracle_matches_reference_script_run():
    """BASELINE configs[0]: oracle/episode.py::OracleSines against what the reference's own sines/train_DKT.py::main()
    computed (tests/golden/make_golden_sines.py ran it unmodified: 50 000 iterations, 500 test tasks)."""
    import dkt_checks
    gold = dkt_checks.check_sines_oracle_against_reference_run()
    assert 0.0 < float(gold["average_mse"]) < 0.1      # the reference's own result on its test tasks
