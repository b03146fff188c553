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
