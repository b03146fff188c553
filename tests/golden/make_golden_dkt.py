"""Golden vectors of the hot path, produced by RUNNING THE REFERENCE'S OWN, UNMODIFIED ``methods/DKT.py`` and
``methods/DKT_regression.py`` (imported from /root/reference, which exists only in the authoring container -- the
fixtures are committed) on its own, unmodified ``backbone.py``.

GPyTorch is not installable offline, so the reference modules are imported on top of oracle/gpytorch_standin: a
self-contained stand-in for the GPyTorch classes those files touch whose arithmetic is INDEPENDENT of oracle/gp.py --
kernels written the way GPyTorch 1.0.1 evaluates them, LAPACK Cholesky / solves in float64 through scipy, analytic
gradient of the Gaussian log-density -- and which this script first validates against scikit-learn's
GaussianProcessRegressor (marginal likelihood, its gradient, predictive mean and standard deviation; linear and RBF
kernels).  So the chain of evidence for every number in these files is

    scikit-learn GPR  <->  stand-in (scipy/LAPACK fp64)  ->  reference DKT.py / DKT_regression.py / backbone.py (unmodified)
                      ->  tests/golden/dkt_*.npz  <-  oracle/episode.py + oracle/gp.py (torch.linalg + autograd)   [tests/test_oracle.py]
                                                  <-  the CUDA path                                                  [tests/test_dkt_gpu.py]

What these fixtures pin: everything the reference itself writes around the GP calls (which images feed which call, the
+-1 target layout, train / eval switching of the backbone, the per-call Adam with its two learning rates, monitoring
predictions conditioned on the pre-update features with post-update hyper-parameters, the ``zip`` truncations,
``correct`` / ``test_loop`` / ``get_logits``, the regression loops) AND the GP arithmetic as an evaluation that shares no
code with oracle/gp.py.  What stays unpinned: a real GPyTorch 1.0.1 install.

``.cuda()`` is patched to the identity so the reference's hard-coded device moves run on the CPU; no reference source is
edited or copied.

Run:  python tests/golden/make_golden_dkt.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

N_WAY, N_SUPPORT, N_QUERY, IMAGE = 3, 2, 3, 84
CLS_KERNELS = ("bncossim", "rbf", "cossim", "linear")


class Recorder:
    """Stands in for the tensorboardX writer (DKT.py:51-55): the reference logs loss and both accuracies through it."""

    def __init__(self):
        self.scalars = {}

    def add_scalar(self, name, value, iteration):
        self.scalars.setdefault(name, []).append(float(value.detach()) if torch.is_tensor(value) else float(value))

    def add_histogram(self, *a, **k):
        pass


def episodes(count, seed, n_way=N_WAY, per_class=N_SUPPORT + N_QUERY, image=IMAGE):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        x = torch.randn(n_way, per_class, 3, image, image, generator=g)
        out.append(x + 0.5 * torch.randn(n_way, 1, 3, 1, 1, generator=g))
    return out


def load_backbone(module, p):
    """oracle parameter dict -> the reference module (ConvBlock registers C / BN twice, backbone.py:115-127)."""
    new = {}
    for k in module.state_dict():
        kk = k
        if kk not in p and ".trunk." in kk[6:]:
            head, tail = kk.split(".trunk.", 1)
            idx, rest = tail.split(".", 1)
            kk = head + (".C." if idx == "0" else ".BN.") + rest
        new[k] = p[kk].clone()
    module.load_state_dict(new)


def gp_perturbation(kernel, n_models):
    """Non-default, per-class-distinct raw hyper-parameters so the fixtures are sensitive to which model gets which."""
    out = {"constant": torch.tensor([0.05 * (c + 1) for c in range(n_models)]),
           "raw_outputscale": torch.tensor([0.1 * c - 0.1 for c in range(n_models)])}
    if kernel in ("rbf", "matern"):
        out["raw_lengthscale"] = torch.tensor([30.0 + 5.0 * c for c in range(n_models)])   # un-normalised D=1600 features
    if kernel == "linear":
        out["raw_variance"] = torch.tensor([-3.0 + 0.3 * c for c in range(n_models)])
    return out


def set_ref_gp(models, pert):
    for c, m in enumerate(models):
        with torch.no_grad():
            m.mean_module.constant.fill_(float(pert["constant"][c]))
            m.covar_module.raw_outputscale.fill_(float(pert["raw_outputscale"][c]))
            if "raw_lengthscale" in pert:
                m.covar_module.base_kernel.raw_lengthscale.fill_(float(pert["raw_lengthscale"][c]))
            if "raw_variance" in pert:
                m.covar_module.base_kernel.raw_variance.fill_(float(pert["raw_variance"][c]))


def ref_gp_state(models, kernel):
    out = {"constant": [], "raw_outputscale": []}
    for m in models:
        out["constant"].append(m.mean_module.constant.detach().reshape(()))
        out["raw_outputscale"].append(m.covar_module.raw_outputscale.detach().reshape(()))
        if kernel in ("rbf", "matern"):
            out.setdefault("raw_lengthscale", []).append(m.covar_module.base_kernel.raw_lengthscale.detach().reshape(()))
        if kernel == "linear":
            out.setdefault("raw_variance", []).append(m.covar_module.base_kernel.raw_variance.detach().reshape(()))
    return {k: torch.stack(v).numpy() for k, v in out.items()}


def classification_case(kernel, ref_backbone, ref_dkt):
    from oracle import backbone as obb
    ref_dkt.kernel_type = kernel                     # ``from configs import kernel_type`` (DKT.py:13): a module global
    torch.manual_seed(0)
    model = ref_dkt.DKT(ref_backbone.Conv4, n_way=N_WAY, n_support=N_SUPPORT)
    p = obb.init_params("Conv4", seed=5, bn_out=(kernel == "bncossim"))
    load_backbone(model.feature, p)
    pert = gp_perturbation(kernel, N_WAY)
    set_ref_gp(model.model.models, pert)
    model.writer = Recorder()

    out = {}
    test_eps = episodes(2, seed=13, per_class=N_SUPPORT + 4)
    model.eval()                                      # test path from the INITIAL weights (device parity without Adam drift)
    out["logits_init"] = np.stack([model.get_logits(x).detach().numpy() for x in test_eps])
    out["correct_init"] = np.array([model.correct(x) for x in test_eps], dtype=np.float64)
    train_eps = episodes(3, seed=11)
    model.train()                                     # train.py:49
    model.train_loop(0, [(x, None) for x in train_eps], None, print_freq=1)
    out["loss"] = np.array(model.writer.scalars["loss"])
    out["acc_support"] = np.array(model.writer.scalars["GP_support_accuracy"])
    out["acc_query"] = np.array(model.writer.scalars["GP_query_accuracy"])
    sd = model.feature.state_dict()
    out["after_conv0_bias"] = sd["trunk.0.C.bias"].numpy().copy()
    out["after_conv3_weight_head"] = sd["trunk.3.C.weight"].numpy().reshape(-1)[:256].copy()
    out["after_bn0_running_mean"] = sd["trunk.0.BN.running_mean"].numpy().copy()
    out["after_bn3_running_var"] = sd["trunk.3.BN.running_var"].numpy().copy()
    if kernel == "bncossim":
        out["after_bn_out_weight_head"] = sd["trunk.bn_out.weight"].numpy()[:256].copy()
        out["after_bn_out_running_mean_head"] = sd["trunk.bn_out.running_mean"].numpy()[:256].copy()
    out["after_param_abs_sum"] = np.float64(sum(float(v.double().abs().sum()) for k, v in sd.items()
                                                if v.is_floating_point() and ".trunk." not in k[6:]))
    for k, v in ref_gp_state(model.model.models, kernel).items():
        out["after_gp_" + k] = v

    # a second train_loop call re-creates Adam (DKT.py:114-115): moments restart
    model.writer = Recorder()
    model.train_loop(1, [(x, None) for x in episodes(1, seed=12)], None, print_freq=1)
    out["loss_second_call"] = np.array(model.writer.scalars["loss"])

    model.eval()                                      # test.py:160
    out["logits"] = np.stack([model.get_logits(x).detach().numpy() for x in test_eps])
    model.writer = None
    acc_mean, acc_std = model.test_loop([(x, None) for x in test_eps], return_std=True)
    out["test_acc_mean"], out["test_acc_std"] = np.float64(acc_mean), np.float64(acc_std)
    out["correct"] = np.array([model.correct(x) for x in test_eps], dtype=np.float64)
    out["correct_adapt"] = np.array(model.correct(test_eps[0], N=3), dtype=np.float64)
    for k, v in ref_gp_state(model.model.models, kernel).items():
        out["adapt_gp_" + k] = v
    out["logits_after_adapt"] = model.get_logits(test_eps[1]).detach().numpy()
    return out


def regression_case(kernel, ref_backbone, ref_reg):
    from oracle import backbone as obb
    ref_reg.kernel_type = kernel
    people, shots = 3, 19
    g = torch.Generator().manual_seed(21)
    batch = torch.rand(people, shots, 3, 100, 100, generator=g)
    labels = torch.rand(people, shots, generator=g) * 2 - 1
    ref_reg.get_batch = lambda who: (batch.clone(), labels.clone())     # the QMUL images are not in the repo
    torch.manual_seed(0)
    model = ref_reg.DKT(ref_backbone.Conv3())
    p = obb.init_params("Conv3", seed=5)
    model.feature_extractor.load_state_dict({k: v.clone() for k, v in p.items()})
    optimizer = torch.optim.Adam([{"params": model.model.parameters(), "lr": 0.001},          # train_regression.py:33-34
                                  {"params": model.feature_extractor.parameters(), "lr": 0.001}])
    losses = []
    mll_forward = model.mll.forward
    model.mll.forward = lambda o, t: (lambda v: (losses.append(float(-v.detach())), v)[1])(mll_forward(o, t))
    model.model.train(); model.feature_extractor.train(); model.likelihood.train()            # train_regression.py:37-39
    model.train_loop(1, optimizer)
    out = {"loss": np.array(losses)}
    out["after_layer1_bias"] = model.feature_extractor.state_dict()["layer1.bias"].numpy().copy()
    out["after_constant"] = model.model.mean_module.constant.detach().numpy().copy()
    out["after_raw_noise"] = model.likelihood.noise_covar.raw_noise.detach().numpy().copy()
    if kernel == "rbf":
        out["after_raw_outputscale"] = model.model.covar_module.raw_outputscale.detach().numpy().reshape(1).copy()
        out["after_raw_lengthscale"] = model.model.covar_module.base_kernel.raw_lengthscale.detach().numpy().reshape(1).copy()
    else:
        out["after_raw_mixture_weights"] = model.model.covar_module.raw_mixture_weights.detach().numpy().copy()
        out["after_raw_mixture_means_head"] = model.model.covar_module.raw_mixture_means.detach().numpy().reshape(4, -1)[:, :64].copy()
    n_support = 5
    ref_reg.test_people = ["a", "b", "c"]             # only its length is read (DKT_regression.py:82)
    np.random.seed(3)                                 # replay of the two draws test_loop makes (DKT_regression.py:69, 82)
    support_ind = list(np.random.choice(list(range(19)), replace=False, size=n_support))
    n = np.random.randint(0, len(ref_reg.test_people) - 1)
    np.random.seed(3)
    mse = model.test_loop(n_support)
    out["test_support_ind"] = np.array(support_ind)
    out["test_person"] = np.int64(n)
    out["test_mse"] = np.float64(float(mse))
    return out


def validate_standin_against_sklearn():
    """The stand-in's marginal likelihood (+ gradient w.r.t. the log hyper-parameters), predictive mean and standard
    deviation against scikit-learn's GaussianProcessRegressor for ScaleKernel(Linear) and ScaleKernel(RBF) + learned
    noise, at float64-rounding level."""
    import gpytorch
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, WhiteKernel
    rs = np.random.RandomState(0)
    n, d, m = 17, 5, 9
    x, xt, y = rs.randn(n, d), rs.randn(m, d), rs.randn(n)
    worst = 0.0
    for kind in ("linear", "rbf"):
        lik = gpytorch.likelihoods.GaussianLikelihood()
        base = gpytorch.kernels.LinearKernel() if kind == "linear" else gpytorch.kernels.RBFKernel()

        class Model(gpytorch.models.ExactGP):
            def __init__(self):
                super().__init__(torch.from_numpy(x), torch.from_numpy(y), lik)
                self.mean_module = gpytorch.means.ConstantMean()
                self.covar_module = gpytorch.kernels.ScaleKernel(base)

            def forward(self, a):
                return gpytorch.distributions.MultivariateNormal(self.mean_module(a), self.covar_module(a))
        model = Model().double()
        lik.double()
        with torch.no_grad():
            model.covar_module.raw_outputscale.fill_(0.3)
            lik.noise_covar.raw_noise.fill_(-0.7)
            if kind == "rbf":
                base.raw_lengthscale.fill_(0.9)
        s = float(model.covar_module.outputscale)
        noise = float(lik.noise)
        if kind == "linear":
            v = float(base.variance)
            sk = ConstantKernel(s * v) * DotProduct(sigma_0=0.0) + WhiteKernel(noise)
        else:
            sk = ConstantKernel(s) * RBF(float(base.lengthscale)) + WhiteKernel(noise)
        gpr = GaussianProcessRegressor(kernel=sk, optimizer=None, alpha=0.0).fit(x, y)
        lml, dlml = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
        model.train(); lik.train()
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
        val = mll(model(*model.train_inputs), model.train_targets) * n
        val.backward()
        worst = max(worst, abs(float(val) - lml) / abs(lml))
        # d/d log(noise) = d/d raw_noise / (d noise/d raw_noise) * noise ; same chain rule for the output scale
        sig = lambda r: 1.0 / (1.0 + np.exp(-r))
        g_noise = float(lik.noise_covar.raw_noise.grad) / sig(-0.7) * noise
        g_scale = float(model.covar_module.raw_outputscale.grad) / sig(0.3) * s
        names = [h.name for h in gpr.kernel_.hyperparameters if not h.fixed]
        g_sk = dict(zip(names, dlml))
        worst = max(worst, abs(g_noise - g_sk["k2__noise_level"]) / abs(g_sk["k2__noise_level"]))
        worst = max(worst, abs(g_scale - g_sk["k1__k1__constant_value"]) / abs(g_sk["k1__k1__constant_value"]))
        if kind == "rbf":
            g_ls = float(base.raw_lengthscale.grad) / sig(0.9) * float(base.lengthscale)
            worst = max(worst, abs(g_ls - g_sk["k1__k2__length_scale"]) / abs(g_sk["k1__k2__length_scale"]))
        model.eval(); lik.eval()
        with torch.no_grad():
            pred = lik(model(torch.from_numpy(xt)))
        mu, sd = gpr.predict(xt, return_std=True)
        worst = max(worst, float(np.abs(pred.mean.numpy() - mu).max() / np.abs(mu).max()))
        worst = max(worst, float(np.abs(pred.stddev.numpy() - sd).max() / np.abs(sd).max()))
    assert worst < 1e-9, worst
    return worst


def main():
    sys.path.insert(0, ROOT)                          # for ``oracle``
    sys.path.insert(0, os.path.join(ROOT, "oracle", "gpytorch_standin"))
    sys.path.insert(0, REF)                           # the reference's backbone / methods / configs / utils / data win
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import types
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))      # data/__init__.py imports feature_loader -> h5py (unused here)
    import backbone as ref_backbone
    import methods.DKT as ref_dkt
    import methods.DKT_regression as ref_reg
    assert ref_dkt.__file__.startswith(REF) and ref_reg.__file__.startswith(REF) and ref_backbone.__file__.startswith(REF)

    print("stand-in vs scikit-learn GaussianProcessRegressor: worst relative difference %.2e" % validate_standin_against_sklearn())
    torch.set_num_threads(1)                          # identical reduction order on any host
    for kernel in CLS_KERNELS:
        out = classification_case(kernel, ref_backbone, ref_dkt)
        np.savez_compressed(os.path.join(HERE, "dkt_cls_%s.npz" % kernel), **out)
        print(kernel, "loss", out["loss"], "acc", out["acc_support"], out["acc_query"], "test", out["test_acc_mean"],
              "correct", out["correct"].tolist(), "adapt", out["correct_adapt"].tolist())
    for kernel in ("rbf", "spectral"):
        out = regression_case(kernel, ref_backbone, ref_reg)
        np.savez_compressed(os.path.join(HERE, "dkt_reg_%s.npz" % kernel), **out)
        print("regression", kernel, "loss", out["loss"], "mse", out["test_mse"], out["test_support_ind"], out["test_person"])


if __name__ == "__main__":
    main()
