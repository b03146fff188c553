"""Parameter containers that mirror the GPyTorch module tree the reference builds in
``DKT.get_model_likelihood_mll`` (methods/DKT.py:58-71, 337-378) so that ``state_dict()`` keys
(``model.models.0.covar_module.raw_outputscale`` ...) and ``model.parameters()`` keep their meaning
with GPyTorch removed.  No arithmetic happens here: the exact-GP math is csrc/gp.cu.
Shapes follow GPyTorch 1.0.1: raw_noise [1], constant [1], raw_outputscale [], raw_lengthscale /
raw_variance / raw_offset [1,1].
"""
import math

import torch
import torch.nn as nn

CLASSIFICATION_KERNELS = ("linear", "rbf", "RBF", "matern", "poli1", "poli2", "cossim", "bncossim")


def inv_softplus(x):
    return x + math.log(-math.expm1(-x))


class _NoiseCovar(nn.Module):
    def __init__(self, noise=None, learn=True):
        super().__init__()
        raw = 0.0 if noise is None else inv_softplus(noise - 1e-4)     # noise = softplus(raw) + 1e-4
        self.raw_noise = nn.Parameter(torch.full((1,), raw), requires_grad=learn)


class GaussianLikelihood(nn.Module):
    def __init__(self, noise=None, learn=True):
        super().__init__()
        self.noise_covar = _NoiseCovar(noise, learn)

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.noise_covar.raw_noise) + 1e-4


class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(1))


class _BaseKernel(nn.Module):
    def __init__(self, kind):
        super().__init__()
        self.kind = kind
        if kind in ("linear", "cossim", "bncossim"):
            frozen = kind != "linear"     # DKT.py:369-370: variance := 1.0, requires_grad False
            self.raw_variance = nn.Parameter(torch.full((1, 1), inv_softplus(1.0) if frozen else 0.0),
                                             requires_grad=not frozen)
        elif kind in ("rbf", "matern"):
            self.raw_lengthscale = nn.Parameter(torch.zeros(1, 1))
        elif kind in ("poli1", "poli2"):
            self.raw_offset = nn.Parameter(torch.zeros(1))

    @property
    def lengthscale(self):
        if hasattr(self, "raw_lengthscale"):
            return torch.nn.functional.softplus(self.raw_lengthscale)
        return None


class ScaleKernel(nn.Module):
    def __init__(self, kind):
        super().__init__()
        self.base_kernel = _BaseKernel(kind)
        self.raw_outputscale = nn.Parameter(torch.zeros(()))

    @property
    def outputscale(self):
        return torch.nn.functional.softplus(self.raw_outputscale)


class SpectralMixtureKernel(nn.Module):
    """gpytorch.kernels.SpectralMixtureKernel(num_mixtures, ard_num_dims) parameter container (no ScaleKernel around
    it, DKT_regression.py:122): raw_mixture_weights [Q], raw_mixture_means / raw_mixture_scales [Q,1,D], all zeros."""

    def __init__(self, num_mixtures=4, ard_num_dims=2916):
        super().__init__()
        self.num_mixtures, self.ard_num_dims = num_mixtures, ard_num_dims
        self.raw_mixture_weights = nn.Parameter(torch.zeros(num_mixtures))
        self.raw_mixture_means = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))
        self.raw_mixture_scales = nn.Parameter(torch.zeros(num_mixtures, 1, ard_num_dims))


class ExactGPLayer(nn.Module):
    """One one-vs-rest exact GP of the classifier (methods/DKT.py:337-378)."""

    def __init__(self, likelihood, kernel="linear"):
        super().__init__()
        k = "rbf" if kernel == "RBF" else kernel
        if k not in ("linear", "rbf", "matern", "poli1", "poli2", "cossim", "bncossim"):
            raise ValueError("[ERROR] the kernel '" + str(kernel) + "' is not supported!")
        self.likelihood = likelihood
        self.mean_module = ConstantMean()
        self.covar_module = ScaleKernel(k)


class LikelihoodList(nn.Module):
    def __init__(self, *likelihoods):
        super().__init__()
        self.likelihoods = nn.ModuleList(likelihoods)


class IndependentModelList(nn.Module):
    """gpytorch.models.IndependentModelList: ``models`` plus the ``likelihood`` alias list GPyTorch registers (the same
    likelihood objects: shared parameters, extra ``model.likelihood.likelihoods.N.*`` keys in ``state_dict()``)."""

    def __init__(self, *models):
        super().__init__()
        self.models = nn.ModuleList(models)
        self.likelihood = LikelihoodList(*[m.likelihood for m in models])


class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class SumMarginalLogLikelihood(nn.Module):
    """gpytorch.mlls.SumMarginalLogLikelihood: besides ``likelihood`` / ``model`` it registers one
    ExactMarginalLogLikelihood(m.likelihood, m) per sub-model under ``mlls`` -- aliases of the same parameters, which is
    where the ``mll.mlls.N.*`` keys of a reference checkpoint come from (train.py:57-65)."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model
        self.mlls = nn.ModuleList([ExactMarginalLogLikelihood(m.likelihood, m) for m in model.models])


def adapt_reference_state(state, own):
    """Map a ``state_dict`` written by the reference (GPyTorch parameter containers, any 1.x version) onto this
    package's containers (train.py:57-65 writes ``{'epoch', 'state': model.state_dict()}``,
    test.py:117-126 / train.py:192-197 load it back).  ``own`` is the receiving module's ``state_dict()``.

    Differences handled (GPyTorch is not installable here, so these follow its published module layout, unverified
    against a real checkpoint -- DESIGN.md section 2):
      * constraint buffers ``*.raw_<p>_constraint.lower_bound / upper_bound`` (Interval registers its bounds as buffers)
        have no counterpart here and are dropped;
      * ``mean_module.raw_constant`` (scalar, GPyTorch >= 1.9) is the older ``mean_module.constant`` ([1]);
      * the same parameter has carried different singleton shapes across versions (``raw_outputscale`` [] / [1],
        ``raw_noise`` [1] / [], ``raw_lengthscale`` [1,1] / [1]): tensors are reshaped when the element count agrees.
    Keys this module does not know are left in place so that ``load_state_dict(strict=True)`` still reports them."""
    out = {}
    for k, v in state.items():
        if k.endswith("_constraint.lower_bound") or k.endswith("_constraint.upper_bound"):
            continue
        if k.endswith("mean_module.raw_constant"):
            k = k[:-len("raw_constant")] + "constant"
        if k in own and hasattr(v, "shape") and tuple(v.shape) != tuple(own[k].shape) and v.numel() == own[k].numel():
            v = v.reshape(own[k].shape)
        out[k] = v
    return out
