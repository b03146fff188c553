"""TEST INFRASTRUCTURE ONLY.  A self-contained stand-in for the part of the ``gpytorch`` API that the reference's
``methods/DKT.py``, ``methods/DKT_regression.py`` and ``sines/train_DKT.py`` touch, so that those files can be imported and
RUN UNMODIFIED in the authoring container (GPyTorch is not installable offline).  tests/golden/make_golden_dkt.py,
make_golden_sines.py and make_golden_state.py execute the reference's own code on top of it.

INDEPENDENT ARITHMETIC.  This file does NOT import oracle/gp.py (tests/test_oracle.py asserts it).  Every formula is
written the way GPyTorch 1.0.1 itself evaluates it (module named per class below), and the linear algebra goes through
LAPACK in float64 via ``scipy.linalg`` (``cholesky`` / ``cho_solve`` / ``solve_triangular``) with a hand-derived
analytic gradient -- whereas oracle/gp.py is torch.linalg in the working precision + autograd.  The golden files
therefore compare two separately written evaluations; make_golden_dkt.py additionally checks this one against
scikit-learn's GaussianProcessRegressor when it runs.  Still not GPyTorch itself: parity of the GP arithmetic against a
real GPyTorch 1.0.1 install remains unpinned.

Module tree / parameter names / shapes follow GPyTorch 1.0.1 (``raw_noise [1]``, ``constant [1]``,
``raw_outputscale []``, ``raw_lengthscale [1,1]``, ``raw_variance [1,1]``, ``raw_offset [1]``, ``raw_mixture_* ``;
``IndependentModelList.likelihood``, ``SumMarginalLogLikelihood.mlls``) so that ``state_dict()`` of the reference modules
built on it has the reference's key set.  Never imported by the product."""
import math
import types
import warnings

import numpy as np
import scipy.linalg as sla
import torch
import torch.nn as nn
import torch.nn.functional as F

__version__ = "1.0.1-standin"
_WORK = torch.float64          # GP arithmetic is carried in float64 and rounded once at the end


# ----------------------------------------------------------------------------- constraints (gpytorch/constraints)
def _inv_softplus(v):
    v = torch.as_tensor(v, dtype=torch.float64)
    return v + torch.log(-torch.expm1(-v))


class Positive(nn.Module):
    """softplus transform, lower bound 0."""
    lower = 0.0

    def transform(self, raw):
        return F.softplus(raw) + self.lower

    def inverse_transform(self, value):
        return _inv_softplus(torch.as_tensor(value, dtype=torch.float64) - self.lower)


class GreaterThan(Positive):
    def __init__(self, lower):
        super().__init__()
        self.lower = float(lower)


# ----------------------------------------------------------------------------- utils/cholesky.py
def psd_safe_cholesky(a, dtype_for_jitter=torch.float32):
    """GPyTorch 1.0.1 ``psd_safe_cholesky``: plain Cholesky; on failure retry with 1e-6, 1e-5, 1e-4 (float32; 1e-8.. for
    float64) ADDED to the diagonal (not cumulative), warning on success; re-raise the first error after three tries."""
    try:
        return sla.cholesky(a, lower=True, check_finite=False)
    except sla.LinAlgError as first:
        if np.isnan(a).any():
            raise RuntimeError("cholesky: matrix contains NaN")
        jitter = 1e-6 if dtype_for_jitter == torch.float32 else 1e-8
        for i in range(3):
            jn = jitter * (10 ** i)
            try:
                l = sla.cholesky(a + jn * np.eye(a.shape[-1]), lower=True, check_finite=False)
                warnings.warn("A not p.d., added jitter of %g to the diagonal" % jn, RuntimeWarning)
                return l
            except sla.LinAlgError:
                continue
        raise RuntimeError("cholesky: matrix not positive definite (%s)" % first)


class _GaussianLogProb(torch.autograd.Function):
    """log N(y; mu, K) with K factored by LAPACK in float64.  d/dK = (alpha alpha^T - K^-1)/2, d/d(y-mu) = -alpha."""

    @staticmethod
    def forward(ctx, diff, covar):
        k = covar.detach().to(torch.float64).numpy()
        r = diff.detach().to(torch.float64).numpy()
        l = psd_safe_cholesky(k)
        alpha = sla.cho_solve((l, True), r, check_finite=False)
        n = r.shape[-1]
        val = -0.5 * (float(r @ alpha) + 2.0 * float(np.log(np.diag(l)).sum()) + n * math.log(2.0 * math.pi))
        ctx.l, ctx.alpha = l, alpha
        ctx.dt = (diff.dtype, covar.dtype)
        return torch.tensor(val, dtype=covar.dtype)

    @staticmethod
    def backward(ctx, g):
        kinv = sla.cho_solve((ctx.l, True), np.eye(ctx.l.shape[0]), check_finite=False)
        dk = 0.5 * (np.outer(ctx.alpha, ctx.alpha) - kinv)
        gg = float(g)
        return (torch.from_numpy(-ctx.alpha * gg).to(ctx.dt[0]), torch.from_numpy(dk * gg).to(ctx.dt[1]))


# ----------------------------------------------------------------------------- distributions/multivariate_normal.py
class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean, self.covariance_matrix = mean, covariance_matrix
        self.loc = mean

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1)

    @property
    def stddev(self):
        return self.variance.sqrt()

    def log_prob(self, value):
        return _GaussianLogProb.apply(value.to(self.mean.dtype) - self.mean, self.covariance_matrix)

    def confidence_region(self):
        std2 = self.stddev * 2
        return self.mean - std2, self.mean + std2


# ----------------------------------------------------------------------------- likelihoods
class HomoskedasticNoise(nn.Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))
        self.raw_noise_constraint = GreaterThan(1e-4)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        v = self.raw_noise_constraint.inverse_transform(value).to(self.raw_noise.dtype)
        self.raw_noise.data.copy_(v.expand(1))


class GaussianLikelihood(nn.Module):
    def __init__(self):
        super().__init__()
        self.noise_covar = HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    def forward(self, dist):
        """p(y | x): the latent covariance plus sigma^2 I."""
        n = dist.mean.shape[-1]
        cov = dist.covariance_matrix
        eye = torch.eye(n, dtype=cov.dtype)
        return MultivariateNormal(dist.mean, cov + self.noise.to(cov.dtype) * eye)


class LikelihoodList(nn.Module):
    def __init__(self, *likelihoods):
        super().__init__()
        self.likelihoods = nn.ModuleList(likelihoods)

    def forward(self, *dists):
        return [lk(d) for lk, d in zip(self.likelihoods, dists)]          # zip: surplus arguments are dropped


# ----------------------------------------------------------------------------- means/constant_mean.py
class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


# ----------------------------------------------------------------------------- kernels
def _sq_dist(x1, x2, x1_eq_x2):
    """kernels/kernel.py ``Distance._sq_dist``: centre both inputs on x1's mean, one matmul of the augmented operands
    [-2 x1, |x1|^2, 1] . [x2, 1, |x2|^2]^T, exact-zero diagonal only when no gradient is required, clamp at 0."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    no_grad = not (x1.requires_grad or x2.requires_grad)
    if x1_eq_x2 and no_grad:
        x2_norm, x2_pad = x1_norm, x1_pad
    else:
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2 and no_grad:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


class Kernel(nn.Module):
    has_lengthscale = False

    def __init__(self):
        super().__init__()
        if self.has_lengthscale:
            self.raw_lengthscale = nn.Parameter(torch.zeros(1, 1))
            self.raw_lengthscale_constraint = Positive()

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    def __call__(self, x1, x2=None):
        same = x2 is None or (x1.shape == x2.shape and torch.equal(x1, x2))
        x2 = x1 if x2 is None else x2
        return self.forward(x1.to(_WORK), x2.to(_WORK), same)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, same):                       # kernels/rbf_kernel.py: exp(-d^2 / 2) of the scaled inputs
        ls = self.lengthscale.to(_WORK)
        return _sq_dist(x1.div(ls), x2.div(ls), same).div(-2).exp()


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu=2.5):
        super().__init__()
        self.nu = nu

    def forward(self, x1, x2, same):                       # kernels/matern_kernel.py
        ls = self.lengthscale.to(_WORK)
        mean = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]
        x1_ = (x1 - mean).div(ls)
        x2_ = (x2 - mean).div(ls)
        distance = _sq_dist(x1_, x2_, same).clamp_min(1e-30).sqrt()
        exp_component = torch.exp(-math.sqrt(self.nu * 2) * distance)
        if self.nu == 0.5:
            constant_component = 1
        elif self.nu == 1.5:
            constant_component = (math.sqrt(3) * distance).add(1)
        else:
            constant_component = (math.sqrt(5) * distance).add(1).add(5.0 / 3.0 * distance ** 2)
        return constant_component * exp_component


class LinearKernel(Kernel):
    def __init__(self):
        super().__init__()
        self.raw_variance = nn.Parameter(torch.zeros(1, 1))
        self.raw_variance_constraint = Positive()

    @property
    def variance(self):
        return self.raw_variance_constraint.transform(self.raw_variance)

    @variance.setter
    def variance(self, value):
        v = self.raw_variance_constraint.inverse_transform(value).to(self.raw_variance.dtype)
        self.raw_variance.data.copy_(v.expand(1, 1))

    def forward(self, x1, x2, same):                       # kernels/linear_kernel.py: (x sqrt(v)) (x' sqrt(v))^T
        sv = self.variance.to(_WORK).sqrt()
        return (x1 * sv).matmul((x2 * sv).transpose(-2, -1))


class PolynomialKernel(Kernel):
    def __init__(self, power):
        super().__init__()
        self.power = power
        self.raw_offset = nn.Parameter(torch.zeros(1))
        self.raw_offset_constraint = Positive()

    @property
    def offset(self):
        return self.raw_offset_constraint.transform(self.raw_offset)

    def forward(self, x1, x2, same):                       # kernels/polynomial_kernel.py
        offset = self.offset.to(_WORK).view(1, 1)
        return (torch.matmul(x1, x2.transpose(-2, -1)) + offset).pow(self.power)


class SpectralMixtureKernel(Kernel):
    def __init__(self, num_mixtures=None, ard_num_dims=1):
        super().__init__()
        self.num_mixtures, self.ard_num_dims = num_mixtures, ard_num_dims
        self.raw_mixture_weights = nn.Parameter(torch.zeros(num_mixtures))
        self.raw_mixture_means = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))
        self.raw_mixture_scales = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))
        self.raw_mixture_weights_constraint = Positive()
        self.raw_mixture_means_constraint = Positive()
        self.raw_mixture_scales_constraint = Positive()

    @property
    def mixture_weights(self):
        return self.raw_mixture_weights_constraint.transform(self.raw_mixture_weights)

    @property
    def mixture_means(self):
        return self.raw_mixture_means_constraint.transform(self.raw_mixture_means)

    @property
    def mixture_scales(self):
        return self.raw_mixture_scales_constraint.transform(self.raw_mixture_scales)

    def forward(self, x1, x2, same):                       # kernels/spectral_mixture_kernel.py
        x1_ = x1.unsqueeze(-3)                             # (k x) n x d
        x2_ = x2.unsqueeze(-3)
        sc, mu = self.mixture_scales.to(_WORK), self.mixture_means.to(_WORK)
        x1_exp, x2_exp = x1_ * sc, x2_ * sc                # scaled BEFORE differencing
        x1_cos, x2_cos = x1_ * mu, x2_ * mu
        exp_term = (x1_exp.unsqueeze(-2) - x2_exp.unsqueeze(-3)).pow(2).mul(-2 * math.pi ** 2)   # k x n x m x d
        cos_term = (x1_cos.unsqueeze(-2) - x2_cos.unsqueeze(-3)).mul(2 * math.pi)
        res = (exp_term.exp() * cos_term.cos()).prod(-1)   # product over dimensions
        w = self.mixture_weights.to(_WORK).unsqueeze(-1).unsqueeze(-1)
        return (res * w).sum(-3)                           # sum over mixtures


class ScaleKernel(Kernel):
    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))
        self.raw_outputscale_constraint = Positive()

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    def forward(self, x1, x2, same):
        return self.base_kernel.forward(x1, x2, same).mul(self.outputscale.to(_WORK))


# ----------------------------------------------------------------------------- models
class ExactGP(nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(i.unsqueeze(-1) if i.dim() == 1 else i for i in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            inputs = (inputs,) if torch.is_tensor(inputs) else tuple(inputs)
            self.train_inputs = tuple(i.unsqueeze(-1) if i.dim() == 1 else i for i in inputs)
        if targets is not None:
            self.train_targets = targets

    def __call__(self, *args):
        inputs = [a.unsqueeze(-1) if a.dim() == 1 else a for a in args]
        if self.training:                                   # models/exact_gp.py: prior at the training inputs
            if not all(torch.equal(t, i) for t, i in zip(self.train_inputs, inputs)):
                raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs)
        # eval: models/exact_prediction_strategies.py DefaultPredictionStrategy -- mean cache K~^-1 (y - mu) (detached),
        # mean* = mu* + K*^T cache, cov* = K** - K*^T K~^-1 K*; LAPACK in float64
        x_tr, x_te = self.train_inputs[0], inputs[0]
        out_dtype = x_te.dtype
        with torch.no_grad():
            prior = self.forward(x_tr)
            k_tt = self.likelihood(prior).covariance_matrix.to(torch.float64).numpy()
            l = psd_safe_cholesky(k_tt)
            r = (self.train_targets.to(torch.float64) - prior.mean.to(torch.float64)).numpy()
            cache = sla.cho_solve((l, True), r, check_finite=False)
            k_st = self.covar_module(x_te, x_tr).to(torch.float64).numpy()              # [M,N]
            mu_te = self.mean_module(x_te).to(torch.float64).numpy()
            mean = mu_te + k_st @ cache
            v = sla.solve_triangular(l, k_st.T, lower=True, check_finite=False)         # [N,M]
            cov = self.covar_module(x_te, x_te).to(torch.float64).numpy() - v.T @ v
        return MultivariateNormal(torch.from_numpy(mean).to(out_dtype), torch.from_numpy(cov).to(out_dtype))


class IndependentModelList(nn.Module):
    def __init__(self, *models):
        super().__init__()
        self.models = nn.ModuleList(models)
        self.likelihood = LikelihoodList(*[m.likelihood for m in models])

    @property
    def train_inputs(self):
        return [m.train_inputs for m in self.models]

    @property
    def train_targets(self):
        return [m.train_targets for m in self.models]

    def __call__(self, *args):
        return [m(*(a if isinstance(a, (tuple, list)) else (a,))) for m, a in zip(self.models, args)]


# ----------------------------------------------------------------------------- mlls
class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output, target):
        res = self.likelihood(output).log_prob(target)
        return res.div(target.size(-1)).to(target.dtype)


class SumMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model
        self.mlls = nn.ModuleList([ExactMarginalLogLikelihood(m.likelihood, m) for m in model.models])

    def forward(self, outputs, targets):
        tot = sum(mll(o, t) for mll, o, t in zip(self.mlls, outputs, targets))
        return tot / len(self.mlls)


class _NoOpSetting:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


likelihoods = types.SimpleNamespace(GaussianLikelihood=GaussianLikelihood, LikelihoodList=LikelihoodList)
models = types.SimpleNamespace(ExactGP=ExactGP, IndependentModelList=IndependentModelList)
mlls = types.SimpleNamespace(ExactMarginalLogLikelihood=ExactMarginalLogLikelihood,
                             SumMarginalLogLikelihood=SumMarginalLogLikelihood)
means = types.SimpleNamespace(ConstantMean=ConstantMean)
kernels = types.SimpleNamespace(ScaleKernel=ScaleKernel, LinearKernel=LinearKernel, RBFKernel=RBFKernel,
                                MaternKernel=MaternKernel, PolynomialKernel=PolynomialKernel,
                                SpectralMixtureKernel=SpectralMixtureKernel)
distributions = types.SimpleNamespace(MultivariateNormal=MultivariateNormal)
constraints = types.SimpleNamespace(Positive=Positive, GreaterThan=GreaterThan)
settings = types.SimpleNamespace(num_likelihood_samples=_NoOpSetting, fast_pred_var=_NoOpSetting)
utils = types.SimpleNamespace(cholesky=types.SimpleNamespace(psd_safe_cholesky=psd_safe_cholesky))
