"""Deep Kernel Transfer, regression (QMUL head pose) -- drop-in for the reference ``methods/DKT_regression.py``
(class ``DKT``: get_model_likelihood_mll 26-37, train_loop 45-64, test_loop 66-97, save/load_checkpoint 99-110;
``ExactGPLayer`` 112-129) with GPyTorch removed: Conv3 features, one exact GP with a ScaleKernel(RBF) base kernel,
constant mean and a LEARNED Gaussian noise; marginal likelihood, its gradients, predictive mean and the
``confidence_region`` variance all run in the sm_100a kernels behind include/dktb200.h.

Like the reference, ``train_loop(epoch, optimizer)`` uses the CALLER's optimizer (built in train_regression.py:33-34
from ``model.model.parameters()`` and ``model.feature_extractor.parameters()``): the kernels write gradients into
``param.grad`` and the caller's ``optimizer.step()`` applies them.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib, configs, gp_modules as gpm
from ..engine import _stream, check_info, GP_JITTER

kernel_type = configs.kernel_type


class ExactGPLayer(nn.Module):
    """Parameter container mirroring DKT_regression.ExactGPLayer (112-129)."""

    def __init__(self, likelihood, kernel="rbf", feat_dim=2916):
        super().__init__()
        self.likelihood = likelihood
        self.mean_module = gpm.ConstantMean()
        if kernel in ("rbf", "RBF"):
            self.covar_module = gpm.ScaleKernel("rbf")
        elif kernel == "spectral":
            self.covar_module = gpm.SpectralMixtureKernel(num_mixtures=4, ard_num_dims=feat_dim)
        else:
            raise ValueError("[ERROR] the kernel '" + str(kernel) + "' is not supported for regression, use 'rbf' or 'spectral'.")


class _Pred:
    """What the reference reads off ``likelihood(model(z))``: ``.mean`` and ``confidence_region()``."""

    def __init__(self, mean, var):
        self.mean, self.variance = mean, var

    def confidence_region(self):
        std2 = self.variance.clamp_min(1e-10).sqrt() * 2.0
        return self.mean - std2, self.mean + std2


class DKT(nn.Module):
    def __init__(self, backbone, kernel=None, lib=None, get_batch=None, feat_dim=2916):
        super(DKT, self).__init__()
        self.feature_extractor = backbone
        self.feat_dim = feat_dim
        self.kernel = kernel if kernel is not None else ("rbf" if kernel_type not in ("rbf", "RBF", "spectral") else kernel_type)
        self._lib = lib
        self._get_batch = get_batch          # injection point for tests; default: the reference's data.qmul_loader
        self.get_model_likelihood_mll()
        self._w = None

    def get_model_likelihood_mll(self, train_x=None, train_y=None):
        likelihood = gpm.GaussianLikelihood()                 # learned noise, default init (DKT_regression.py:29)
        model = ExactGPLayer(likelihood, kernel=self.kernel, feat_dim=self.feat_dim)
        self.model = model
        self.likelihood = likelihood
        self.mll = gpm.ExactMarginalLogLikelihood(likelihood, model)     # DKT_regression.py:34
        self.mse = nn.MSELoss()
        return self.model, self.likelihood, self.mll

    def set_forward(self, x, is_feature=False):
        pass

    def set_forward_loss(self, x):
        pass

    # ------------------------------------------------------------------ kernels
    def _dev(self):
        return next(self.feature_extractor.parameters()).device

    def _lib_get(self):
        dev = self._dev()
        if dev.type != "cuda" and self._lib is None:
            raise RuntimeError("dktb200 has no CPU path: call .cuda() on the model first")
        self.lib = self._lib or _lib.load()
        return self.lib

    def _hyper(self):
        m = self.model
        if self.kernel == "spectral":
            return (None, m.mean_module.constant.data.view(1), self.likelihood.noise_covar.raw_noise.data.view(1), None)
        return (m.covar_module.raw_outputscale.data.view(1), m.mean_module.constant.data.view(1),
                self.likelihood.noise_covar.raw_noise.data.view(1), m.covar_module.base_kernel.raw_lengthscale.data.view(1))

    def _spectral(self):
        k = self.model.covar_module
        Q = k.num_mixtures
        return k.raw_mixture_weights.data, k.raw_mixture_means.data.view(Q, -1), k.raw_mixture_scales.data.view(Q, -1), Q

    def _work(self, N, M, D, dev):
        key = (N, M, D, str(dev))
        if self._w is None or self._w["key"] != key:
            f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            self._w = dict(key=key, xc=f(1, N, D), sq=f(1, N), gram=f(1, N, N), kb=f(1, 1, N, N), alpha=f(1, 1, N),
                           linv=f(1, 1, N, N), lt=f(1, 1), info=torch.zeros(1, 1, device=dev, dtype=torch.int32),
                           dk=f(1, 1, N, N), dh=f(1, 1, 3), dg=f(1, N, N), dparam=f(1), ksc=f(N), dz=f(1, N, D),
                           xt=f(1, M, D), sqt=f(1, M), gx=f(1, M, N), kx=f(1, 1, M, N), mean=f(1, 1, M), var=f(1, 1, M),
                           kss=f(1, 1, M), loss=f(1), hyper=f(1, 3))
            if self.kernel == "spectral":
                Q = self.model.covar_module.num_mixtures
                self._w.update(ec=f(1, Q, N, N), dw=f(Q), dmu=f(Q, D), dv=f(Q, D))
        return self._w

    def _fit(self, z, y, want_grad):
        """z [N,D] features, y [N] targets: K~ -> Cholesky -> alpha / loss (/ gradients).  Returns the workspace."""
        lib, dev = self.lib, z.device
        N, D = z.shape
        st = _stream(dev)
        w = self._work(N, self._w["key"][1] if self._w else N, D, dev) if self._w and self._w["key"][0] == N and self._w["key"][2] == D \
            else self._work(N, N, D, dev)
        ros, cst, rn, rl = self._hyper()
        zz = z.contiguous().view(1, N, D)
        if self.kernel == "spectral":
            rw, rmu, rv, Q = self._spectral()
            w["xc"].copy_(zz)
            lib.spectral_fwd(w["xc"], w["xc"], rw, rmu, rv, w["kb"], w["ec"], 1, N, N, D, Q, 36, self._P, st)
        else:
            w["xc"].copy_(zz)
            lib.sqdist(w["xc"], w["xc"], w["gram"], 1, N, N, D, st)          # w["gram"] holds squared distances
            lib.kernel_fwd(1, None, w["gram"], rl, w["kb"], 1, 1, N, N, st)
        lib.gp_fit(w["kb"], N * N, y.contiguous().view(1, 1, N), N, ros, cst, rn, w["alpha"], w["linv"], w["lt"], w["info"],
                   w["dk"] if want_grad else None, w["dh"] if want_grad else None, 1.0, GP_JITTER, 1, 1, N, st)
        lib.gp_reduce(w["lt"], w["dh"] if want_grad else None, w["loss"], w["hyper"] if want_grad else None, 1, 1, st)
        return w

    @staticmethod
    def _set_grad(p, g):
        g = g.reshape(p.shape)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)

    def train_step(self, inputs, labels):
        """One person (DKT_regression.py:48-57, minus optimizer.zero_grad/step): returns the loss (device scalar) and
        leaves the gradients in ``param.grad``."""
        lib = self._lib_get()
        fe = self.feature_extractor
        eng = fe.engine(inputs.device, lib)
        ws = [l.weight.data for l in fe.layers()]
        bs = [l.bias.data for l in fe.layers()]
        z = eng.forward(inputs.contiguous().float(), ws, bs)
        N, D = z.shape
        self._P = eng.P
        w = self._fit(z, labels.float(), want_grad=True)
        st = _stream(inputs.device)
        ros, cst, rn, rl = self._hyper()
        if self.kernel == "spectral":
            rw, rmu, rv, Q = self._spectral()
            lib.spectral_bwd(w["xc"], rw, rmu, rv, w["dk"], w["ec"], w["dw"], w["dmu"], w["dv"], w["dz"], 1, N, D, Q, 36,
                             self._P, st)
        else:
            lib.kernel_bwd(1, None, w["gram"], rl, w["dk"], w["dg"], w["dparam"], w["ksc"], 1, 1, N, st)
            lib.gram_bwd(w["dg"], w["xc"], w["dz"], 1, 1, N, D, 1.0, st)
        gw = [torch.empty_like(t) for t in ws]
        gb = [torch.empty_like(t) for t in bs]
        eng.backward(w["dz"].view(N, D), ws, gw, gb)
        for l, a, b in zip(fe.layers(), gw, gb):
            self._set_grad(l.weight, a)
            self._set_grad(l.bias, b)
        m = self.model
        self._set_grad(m.mean_module.constant, w["hyper"][0, 1])
        self._set_grad(self.likelihood.noise_covar.raw_noise, w["hyper"][0, 2])
        if self.kernel == "spectral":
            self._set_grad(m.covar_module.raw_mixture_weights, w["dw"])
            self._set_grad(m.covar_module.raw_mixture_means, w["dmu"])
            self._set_grad(m.covar_module.raw_mixture_scales, w["dv"])
        else:
            self._set_grad(m.covar_module.raw_outputscale, w["hyper"][0, 0])
            self._set_grad(m.covar_module.base_kernel.raw_lengthscale, w["dparam"])
        self._last_info = w["info"]
        return w["loss"][0].clone()

    def train_loop(self, epoch, optimizer):
        get_batch, people = self._batch_source("train")
        batch, batch_labels = get_batch(people)
        dev = self._dev()
        batch, batch_labels = batch.to(dev), batch_labels.to(dev)
        for inputs, labels in zip(batch, batch_labels):
            optimizer.zero_grad()
            loss = self.train_step(inputs, labels)
            optimizer.step()
            # `predictions.mean` of the train-mode prior is a view of the constant-mean parameter (DKT_regression.py:58)
            mse = ((self.model.mean_module.constant.data.view(()) - labels.float()) ** 2).mean()
            if epoch % 10 == 0:
                check_info(self._last_info)
                print('[%d] - Loss: %.3f  MSE: %.3f noise: %.3f' % (epoch, loss.item(), mse.item(),
                                                                  self.likelihood.noise.item()))

    def _batch_source(self, which):
        if self._get_batch is not None:
            return self._get_batch, which
        from data.qmul_loader import get_batch, train_people, test_people   # the reference's own loader
        return get_batch, (train_people if which == "train" else test_people)

    def predict(self, x_support, y_support, x_query):
        """Fit on the support set, predict the query images: returns a ``_Pred`` (mean, variance incl. noise)."""
        lib = self._lib_get()
        fe = self.feature_extractor
        eng = fe.engine(x_support.device, lib)
        ws = [l.weight.data for l in fe.layers()]
        bs = [l.bias.data for l in fe.layers()]
        st = _stream(x_support.device)
        z_s = eng.forward(x_support.contiguous().float(), ws, bs).clone()
        z_q = eng.forward(x_query.contiguous().float(), ws, bs).clone()
        N, D = z_s.shape
        M = z_q.shape[0]
        self._P = eng.P
        self._w = None
        self._work(N, M, D, z_s.device)
        w = self._fit(z_s, y_support.float(), want_grad=False)
        ros, cst, rn, rl = self._hyper()
        if self.kernel == "spectral":
            rw, rmu, rv, Q = self._spectral()
            lib.spectral_fwd(z_q.view(1, M, D), w["xc"], rw, rmu, rv, w["kx"], None, 1, M, N, D, Q, 36, self._P, st)
            w["kss"].copy_(torch.nn.functional.softplus(rw).sum().expand(1, 1, M))   # k(x*,x*) = sum_q w_q
        else:
            lib.sqdist(z_q.view(1, M, D), w["xc"], w["gx"], 1, M, N, D, st)
            lib.kernel_fwd(1, None, w["gx"], rl, w["kx"], 1, 1, M, N, st)
            w["kss"].fill_(1.0)          # k_rbf(x*, x*) = 1
        lib.gp_predict(w["kx"], M * N, w["alpha"], ros, cst, w["mean"], None, 1, 1, M, N, st)
        lib.gp_predict_var(w["kx"], M * N, w["kss"], M, w["linv"], ros, rn, w["var"], 1, 1, M, N, st)
        self._last_info = w["info"]
        return _Pred(w["mean"].view(M).clone(), w["var"].view(M).clone())

    def test_loop(self, n_support, optimizer=None):      # no optimizer needed for GP
        get_batch, people = self._batch_source("test")
        inputs, targets = get_batch(people)
        support_ind = list(np.random.choice(list(range(19)), replace=False, size=n_support))
        dev = self._dev()
        x_all, y_all = inputs.to(dev), targets.to(dev)
        x_support = inputs[:, support_ind, :, :, :].to(dev)
        y_support = targets[:, support_ind].to(dev)
        n = np.random.randint(0, inputs.shape[0] - 1)     # a random test person
        self.model.eval()
        self.feature_extractor.eval()
        self.likelihood.eval()
        pred = self.predict(x_support[n], y_support[n], x_all[n])
        check_info(self._last_info)
        lower, upper = pred.confidence_region()            # 2 standard deviations above and below the mean
        mse = self.mse(pred.mean, y_all[n].float())
        return mse

    def save_checkpoint(self, checkpoint):
        torch.save({'gp': self.model.state_dict(), 'likelihood': self.likelihood.state_dict(),
                    'net': self.feature_extractor.state_dict()}, checkpoint)

    def load_checkpoint(self, checkpoint):
        ckpt = torch.load(checkpoint)
        self.model.load_state_dict(gpm.adapt_reference_state(ckpt['gp'], self.model.state_dict()))
        self.likelihood.load_state_dict(gpm.adapt_reference_state(ckpt['likelihood'], self.likelihood.state_dict()))
        self.feature_extractor.load_state_dict(ckpt['net'])
