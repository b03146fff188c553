"""TEST INFRASTRUCTURE ONLY.  A minimal stand-in for the part of the ``gpytorch`` API that the reference's
``methods/DKT.py`` and ``methods/DKT_regression.py`` touch, so that those files can be imported and RUN UNMODIFIED in the
authoring container (GPyTorch is not installable offline).  It exists for one purpose: tests/golden/make_golden_dkt.py
executes the reference's own ``train_loop`` / ``correct`` / ``get_logits`` code on top of it to pin the CONTROL FLOW that
oracle/episode.py restates (which features feed which GP call, train / eval switching, target layout, the optimiser
wiring, the monitoring predictions with pre-update features and post-update hyper-parameters, the zip truncations).

The GP ARITHMETIC here is oracle/gp.py's restatement of GPyTorch 1.0.1 semantics -- it is NOT
GPyTorch, so parity of the arithmetic against GPyTorch itself stays unpinned.  Module / parameter
names and shapes follow GPyTorch 1.0.1.  Never imported by the product."""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import gp as ogp


def _inv_softplus(v):
    return ogp.inv_softplus(torch.as_tensor(v, dtype=torch.float64)).to(torch.float32)


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean, self.covariance_matrix = mean, covariance_matrix

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1)

    def log_prob(self, y):
        import math
        n = y.shape[-1]
        l = torch.linalg.cholesky(self.covariance_matrix)
        r = (y - self.mean).unsqueeze(-1)
        alpha = torch.cholesky_solve(r, l)
        return -0.5 * ((r * alpha).sum() + 2.0 * torch.log(torch.diagonal(l)).sum() + n * math.log(2.0 * math.pi))

    def confidence_region(self):
        std2 = self.variance.sqrt() * 2
        return self.mean - std2, self.mean + std2


class _HomoskedasticNoise(nn.Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))

    @property
    def noise(self):
        return F.softplus(self.raw_noise) + ogp.NOISE_LOWER_BOUND

    @noise.setter
    def noise(self, value):
        self.raw_noise.data.copy_(_inv_softplus(torch.as_tensor(value, dtype=torch.float32) - ogp.NOISE_LOWER_BOUND).expand(1))


class GaussianLikelihood(nn.Module):
    def __init__(self):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    def forward(self, dist):
        n = dist.mean.shape[-1]
        return MultivariateNormal(dist.mean, dist.covariance_matrix + self.noise * torch.eye(n, dtype=dist.mean.dtype))


class LikelihoodList(nn.Module):
    def __init__(self, *likelihoods):
        super().__init__()
        self.likelihoods = nn.ModuleList(likelihoods)

    def forward(self, *dists):
        return [lk(d) for lk, d in zip(self.likelihoods, dists)]          # zip: surplus arguments are dropped


class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class _Kernel(nn.Module):
    kind = None
    has_lengthscale = False

    def _params(self):
        return {}

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale) if self.has_lengthscale else None

    def forward(self, x1, x2=None):
        x2 = x1 if x2 is None else x2
        return ogp.base_kernel(self.kind, x1, x2, self._params())


class LinearKernel(_Kernel):
    kind = "linear"

    def __init__(self):
        super().__init__()
        self.raw_variance = nn.Parameter(torch.zeros(1, 1))

    @property
    def variance(self):
        return F.softplus(self.raw_variance)

    @variance.setter
    def variance(self, value):
        self.raw_variance.data.copy_(_inv_softplus(value).expand(1, 1))

    def _params(self):
        return {"raw_variance": [self.raw_variance.view(())]}


class RBFKernel(_Kernel):
    kind = "rbf"
    has_lengthscale = True

    def __init__(self):
        super().__init__()
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, 1))

    def _params(self):
        return {"raw_lengthscale": [self.raw_lengthscale.view(())]}


class MaternKernel(RBFKernel):
    kind = "matern"          # nu = 2.5, GPyTorch's default


class PolynomialKernel(_Kernel):
    def __init__(self, power):
        super().__init__()
        self.kind = {1: "poli1", 2: "poli2"}[power]
        self.raw_offset = nn.Parameter(torch.zeros(1))

    def _params(self):
        return {"raw_offset": [self.raw_offset.view(())]}


class SpectralMixtureKernel(_Kernel):
    kind = "spectral"

    def __init__(self, num_mixtures=4, ard_num_dims=1):
        super().__init__()
        self.raw_mixture_weights = nn.Parameter(torch.zeros(num_mixtures))
        self.raw_mixture_means = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))
        self.raw_mixture_scales = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))

    def _params(self):
        return {"raw_mixture_weights": self.raw_mixture_weights, "raw_mixture_means": self.raw_mixture_means,
                "raw_mixture_scales": self.raw_mixture_scales}


class ScaleKernel(_Kernel):
    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self):
        return F.softplus(self.raw_outputscale)

    def forward(self, x1, x2=None):
        return self.outputscale * self.base_kernel(x1, x2)


class ExactGP(nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,) if torch.is_tensor(train_inputs) else tuple(train_inputs)
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
        if self.training:
            if not all(torch.equal(t, i) for t, i in zip(self.train_inputs, inputs)):
                raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs)
        # eval: exact predictive distribution conditioned on the stored training data (mean cache detached)
        x_tr, x_te = self.train_inputs[0], inputs[0]
        k_tt = self.covar_module(x_tr, x_tr)
        n = x_tr.shape[-2]
        l = torch.linalg.cholesky(k_tt + self.likelihood.noise * torch.eye(n, dtype=x_tr.dtype))
        r = (self.train_targets - self.mean_module(x_tr)).unsqueeze(-1)
        alpha = torch.cholesky_solve(r, l).detach()
        k_st = self.covar_module(x_te, x_tr)
        mean = self.mean_module(x_te) + (k_st @ alpha).squeeze(-1)
        v = torch.linalg.solve_triangular(l, k_st.transpose(-1, -2), upper=False)
        cov = self.covar_module(x_te, x_te) - v.transpose(-1, -2) @ v
        return MultivariateNormal(mean, cov)


class IndependentModelList(nn.Module):
    def __init__(self, *models):
        super().__init__()
        self.models = nn.ModuleList(models)

    @property
    def train_inputs(self):
        return [m.train_inputs for m in self.models]

    @property
    def train_targets(self):
        return [m.train_targets for m in self.models]

    def __call__(self, *args):
        return [m(*(a if isinstance(a, (tuple, list)) else (a,))) for m, a in zip(self.models, args)]


class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output, target):
        return self.likelihood(output).log_prob(target) / target.size(-1)


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
settings = types.SimpleNamespace(num_likelihood_samples=_NoOpSetting, fast_pred_var=_NoOpSetting)
