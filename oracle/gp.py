"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Exact-GP arithmetic of the DKT path.

Restates what the reference obtains from GPyTorch (un-vendored third-party dependency,
README.md:24-37: ">=0.3.5", known-good 1.0.1) at the call sites
  methods/DKT.py:58-71 (model/likelihood/mll construction), 161-162 (prior + -mll),
  177/187/265/330 (predictive mean), 337-378 (ExactGPLayer kernels),
  methods/DKT_regression.py:29-34, 52-54, 90-93, 112-129,
  sines/train_DKT.py:126-143, 171-180, 199-229.
PARITY UNPINNED against GPyTorch itself (not installable offline); cross-checked against
scikit-learn GaussianProcessRegressor (tests/test_oracle.py).

Everything is plain differentiable torch so autograd provides the reference gradients
("what autograd-through-Cholesky yields").
"""
import math

import torch
import torch.nn.functional as F

KERNELS = ("linear", "rbf", "matern", "poli1", "poli2", "cossim", "bncossim", "spectral")
NOISE_LOWER_BOUND = 1e-4  # GaussianLikelihood: noise = softplus(raw_noise) + 1e-4  (GreaterThan(1e-4))


def inv_softplus(x):
    x = torch.as_tensor(x, dtype=torch.float64)
    return x + torch.log(-torch.expm1(-x))


def classification_raw_noise(dtype=torch.float32):
    """DKT.py:346-347 -- ``likelihood.noise_covar.noise = 0.1`` with the gradient frozen."""
    return inv_softplus(0.1 - NOISE_LOWER_BOUND).to(dtype)


def default_gp_params(kernel, n_models, feat_dim=None, dtype=torch.float32, classification=True):
    """Raw (pre-softplus) GP hyper-parameters at their GPyTorch initial values (all zeros)."""
    p = {}
    z = lambda *s: torch.zeros(*s, dtype=dtype)
    p["constant"] = z(n_models)                      # ConstantMean
    if classification:
        p["raw_noise"] = classification_raw_noise(dtype).repeat(n_models)
    else:
        p["raw_noise"] = z(n_models)                  # learned, DKT_regression.py:29
    if kernel == "spectral":                          # no ScaleKernel around it (DKT_regression.py:122)
        q = 4
        p["raw_mixture_weights"] = z(q)
        p["raw_mixture_means"] = z(q, 1, feat_dim)
        p["raw_mixture_scales"] = z(q, 1, feat_dim)
        return p
    p["raw_outputscale"] = z(n_models)               # ScaleKernel
    if kernel in ("rbf", "matern"):
        p["raw_lengthscale"] = z(n_models)
    if kernel in ("linear", "cossim", "bncossim"):
        # cossim/bncossim: variance := 1.0, frozen (DKT.py:369-370)
        v = z(n_models) if kernel == "linear" else inv_softplus(1.0).to(dtype).repeat(n_models)
        p["raw_variance"] = v
    if kernel in ("poli1", "poli2"):
        p["raw_offset"] = z(n_models)
    return p


def trainable_gp_names(kernel, classification=True):
    names = ["constant"]
    if kernel == "spectral":
        names += ["raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"]
    else:
        names.append("raw_outputscale")
        if kernel in ("rbf", "matern"):
            names.append("raw_lengthscale")
        if kernel == "linear":
            names.append("raw_variance")
        if kernel in ("poli1", "poli2"):
            names.append("raw_offset")
    if not classification:
        names.append("raw_noise")
    return names


def _sq_dist(a, b):
    a2 = (a * a).sum(-1, keepdim=True)
    b2 = (b * b).sum(-1, keepdim=True)
    d = a2 + b2.transpose(-1, -2) - 2.0 * (a @ b.transpose(-1, -2))
    return d.clamp_min(0.0)


def base_kernel(kernel, x1, x2, p, c=0):
    """Un-scaled kernel matrix K_base(x1, x2) of sub-model ``c`` (DKT.py:352-372)."""
    if kernel in ("linear", "cossim", "bncossim"):
        v = F.softplus(p["raw_variance"][c])
        return v * (x1 @ x2.transpose(-1, -2))
    if kernel in ("rbf", "matern"):
        ls = F.softplus(p["raw_lengthscale"][c])
        mean = x1.mean(-2, keepdim=True)              # GPyTorch centres both inputs on x1's mean
        a = (x1 - mean) / ls
        b = (x2 - mean) / ls
        d2 = _sq_dist(a, b)
        if kernel == "rbf":
            return torch.exp(-0.5 * d2)
        d = d2.clamp_min(1e-30).sqrt()
        s5 = math.sqrt(5.0)
        return (s5 * d + 1.0 + (5.0 / 3.0) * d * d) * torch.exp(-s5 * d)
    if kernel in ("poli1", "poli2"):
        off = F.softplus(p["raw_offset"][c])
        k = x1 @ x2.transpose(-1, -2) + off
        return k if kernel == "poli1" else k * k
    if kernel == "spectral":
        w = F.softplus(p["raw_mixture_weights"])              # [Q]
        mu = F.softplus(p["raw_mixture_means"])               # [Q,1,D]
        sc = F.softplus(p["raw_mixture_scales"])              # [Q,1,D]
        tau = x1.unsqueeze(-2) - x2.unsqueeze(-3)             # [N,M,D]
        tau = tau.unsqueeze(0)                                # [1,N,M,D]
        e = torch.exp(-2.0 * math.pi ** 2 * (tau * sc.unsqueeze(1)) ** 2)
        cs = torch.cos(2.0 * math.pi * tau * mu.unsqueeze(1))
        comp = (e * cs).prod(-1)                              # [Q,N,M]
        return (w.view(-1, 1, 1) * comp).sum(0)
    raise ValueError("[ERROR] the kernel '" + str(kernel) + "' is not supported!")


def outputscale(kernel, p, c=0):
    if kernel == "spectral":
        return torch.ones((), dtype=p["constant"].dtype)
    return F.softplus(p["raw_outputscale"][c])


def noise(p, c=0):
    return F.softplus(p["raw_noise"][c]) + NOISE_LOWER_BOUND


def train_covar(kernel, x, p, c=0):
    """K~_c = s_c K_base(x,x) + sigma^2 I."""
    k = outputscale(kernel, p, c) * base_kernel(kernel, x, x, p, c)
    return k + noise(p, c) * torch.eye(x.shape[-2], dtype=x.dtype)


def log_marginal(kernel, x, y, p, c=0):
    """log p(y | x) of sub-model c (MultivariateNormal.log_prob via dense Cholesky)."""
    n = x.shape[-2]
    kt = train_covar(kernel, x, p, c)
    l = torch.linalg.cholesky(kt)
    r = (y - p["constant"][c]).unsqueeze(-1)
    alpha = torch.cholesky_solve(r, l)
    quad = (r * alpha).sum()
    logdet = 2.0 * torch.log(torch.diagonal(l)).sum()
    return -0.5 * (quad + logdet + n * math.log(2.0 * math.pi))


def mll_loss(kernel, x, targets, p):
    """loss = -SumMarginalLogLikelihood (DKT.py:162): mean over models of (log p / N).

    ``targets``: [C, N].  For a single GP (regression) pass targets of shape [1, N]
    (ExactMarginalLogLikelihood divides by N; DKT_regression.py:54, sines/train_DKT.py:178).
    """
    n = x.shape[-2]
    c_models = targets.shape[0]
    tot = 0.0
    for c in range(c_models):
        tot = tot + log_marginal(kernel, x, targets[c], p, c) / n
    return -(tot / c_models)


def predict(kernel, x_train, targets, x_test, p, want_var=False):
    """Predictive mean (and likelihood-augmented variance) of every sub-model.

    mean_c = m_c + s_c K_base(x*,X) alpha_c,  alpha_c = K~_c^{-1}(y_c - m_c)  (detached caches).
    var_c  = diag(s_c K_base(x*,x*) - K*^T K~^{-1} K*) + sigma^2    (``likelihood(model(x))``).
    Returns mean [C, M] (and var [C, M]).
    """
    means, variances = [], []
    for c in range(targets.shape[0]):
        kt = train_covar(kernel, x_train, p, c)
        l = torch.linalg.cholesky(kt)
        r = (targets[c] - p["constant"][c]).unsqueeze(-1)
        alpha = torch.cholesky_solve(r, l)
        s = outputscale(kernel, p, c)
        kx = s * base_kernel(kernel, x_test, x_train, p, c)          # [M,N]
        means.append(p["constant"][c] + (kx @ alpha).squeeze(-1))
        if want_var:
            kss = s * base_kernel(kernel, x_test, x_test, p, c)
            v = torch.linalg.solve_triangular(l, kx.transpose(-1, -2), upper=False)   # [N,M]
            variances.append(torch.diagonal(kss) - (v * v).sum(0) + noise(p, c))
    mean = torch.stack(means, 0)
    if want_var:
        return mean, torch.stack(variances, 0)
    return mean


def classify(mean):
    """DKT.py:266-269: sigmoid(mean) per model, vstack, argmax over models (first index on ties)."""
    return torch.sigmoid(mean).argmax(0)
