"""1-D sine regression with DKT (reference sines/train_DKT.py: Task_Distribution 83-111, Feature 113-124,
ExactGPModel 126-143, training loop 171-180, test 199-229) on the dktb200 kernels: MLP 1 -> 40 -> 40 (ReLU) as 1x1
NHWC convolutions, spectral-mixture GP (Q = 4, ARD over the 40 features, no ScaleKernel), learned Gaussian noise,
Adam(lr 1e-3) on one flat parameter buffer.  BASELINE.json config #1 (the reference script itself is CPU-only).
No plotting (matplotlib / seaborn are not part of the path)."""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, gp_modules as gpm
from .engine import _stream, GP_JITTER
from .flat import FlatPack


class Sine_Task:
    def __init__(self, amplitude, phase, xmin, xmax):
        self.amplitude, self.phase, self.xmin, self.xmax = amplitude, phase, xmin, xmax

    def true_function(self, x):
        return self.amplitude * np.sin(self.phase + x)

    def sample_data(self, size=1, noise=0.0, sort=False):
        x = np.random.uniform(self.xmin, self.xmax, size)
        if sort:
            x = np.sort(x)
        y = self.true_function(x)
        if noise > 0:
            y += np.random.normal(loc=0.0, scale=noise, size=y.shape)
        return torch.tensor(x, dtype=torch.float).unsqueeze(1), torch.tensor(y, dtype=torch.float)


class Task_Distribution:
    def __init__(self, amplitude_min=0.1, amplitude_max=5.0, phase_min=0.0, phase_max=np.pi, x_min=-5.0, x_max=5.0,
                 family="sine"):
        self.amplitude_min, self.amplitude_max = amplitude_min, amplitude_max
        self.phase_min, self.phase_max, self.x_min, self.x_max = phase_min, phase_max, x_min, x_max
        self.family = family

    def sample_task(self):
        amplitude = np.random.uniform(self.amplitude_min, self.amplitude_max)
        phase = np.random.uniform(self.phase_min, self.phase_max)
        return Sine_Task(amplitude, phase, self.x_min, self.x_max)


class Feature(nn.Module):
    """Linear(1,40) -> ReLU -> Linear(40,40) -> ReLU (sines/train_DKT.py:113-124).  Parameter container."""

    def __init__(self):
        super().__init__()
        self.layer1 = nn.Linear(1, 40)
        self.layer2 = nn.Linear(40, 40)


class SinesDKT(nn.Module):
    D, Q = 40, 4

    def __init__(self, lib=None, lr=1e-3):
        super().__init__()
        self.net = Feature()
        self.likelihood = gpm.GaussianLikelihood()
        self.mean_module = gpm.ConstantMean()
        self.covar_module = gpm.SpectralMixtureKernel(num_mixtures=self.Q, ard_num_dims=self.D)
        self._lib, self.lr = lib, lr
        self._pack = None

    # ------------------------------------------------------------------ plumbing
    def _ensure(self):
        dev = self.net.layer1.weight.device
        if self._pack is not None and self._pack.intact():
            return
        if dev.type != "cuda" and self._lib is None:
            raise RuntimeError("dktb200 has no CPU path: call .cuda() on the model first")
        self.lib = self._lib or _lib.load()
        self._pack = FlatPack([(n, p, 4) for n, p in self.named_parameters()], dev)
        n = self._pack.numel
        self._adam = {"m": torch.zeros(n, device=dev), "v": torch.zeros(n, device=dev), "step": 0}
        self._w = None

    def _features(self, x, keep):
        """x [N,1] -> z [N,40]; ``keep`` stores the hidden activation for the backward pass."""
        lib, st = self.lib, _stream(x.device)
        N = x.shape[0]
        h = torch.empty(N, 1, 1, 40, device=x.device)
        z = torch.empty(N, 1, 1, 40, device=x.device)
        l1, l2 = self.net.layer1, self.net.layer2
        lib.conv2d_fwd(x.contiguous().view(N, 1, 1, 1), l1.weight.data.view(40, 1, 1, 1), l1.bias.data, h, N, 1, 1, 1, 40, 1,
                       1, 1, 0, 1, 1, st)
        lib.conv2d_fwd(h, l2.weight.data.view(40, 40, 1, 1), l2.bias.data, z, N, 1, 1, 40, 40, 1, 1, 1, 0, 1, 1, st)
        if keep:
            self._saved = (x.contiguous().view(N, 1, 1, 1), h, z)
        return z.view(N, 40)

    def _spectral(self):
        k = self.covar_module
        return k.raw_mixture_weights.data, k.raw_mixture_means.data.view(self.Q, -1), k.raw_mixture_scales.data.view(self.Q, -1)

    def _fit(self, z, y, want_grad):
        lib, st, dev = self.lib, _stream(z.device), z.device
        N, D, Q = z.shape[0], self.D, self.Q
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        w = dict(kb=f(1, 1, N, N), ec=f(1, Q, N, N), alpha=f(1, 1, N), linv=f(1, 1, N, N), lt=f(1, 1),
                 info=torch.zeros(1, 1, device=dev, dtype=torch.int32), dk=f(1, 1, N, N), dh=f(1, 1, 3), loss=f(1),
                 hyper=f(1, 3), x=z.contiguous().view(1, N, D))
        rw, rmu, rv = self._spectral()
        lib.spectral_fwd(w["x"], w["x"], rw, rmu, rv, w["kb"], w["ec"], 1, N, N, D, Q, D, 1, st)
        lib.gp_fit(w["kb"], N * N, y.contiguous().view(1, 1, N), N, None, self.mean_module.constant.data.view(1),
                   self.likelihood.noise_covar.raw_noise.data.view(1), w["alpha"], w["linv"], w["lt"], w["info"],
                   w["dk"] if want_grad else None, w["dh"] if want_grad else None, 1.0, GP_JITTER, 1, 1, N, st)
        lib.gp_reduce(w["lt"], w["dh"] if want_grad else None, w["loss"], w["hyper"] if want_grad else None, 1, 1, st)
        return w

    # ------------------------------------------------------------------ train / predict
    def train_step(self, x, y):
        """One iteration of sines/train_DKT.py:171-180; returns the loss (device scalar)."""
        self._ensure()
        lib, st = self.lib, _stream(x.device)
        z = self._features(x.float(), keep=True)
        N = z.shape[0]
        w = self._fit(z, y.float(), want_grad=True)
        rw, rmu, rv = self._spectral()
        k = self.covar_module
        dz = torch.empty(1, N, self.D, device=x.device)
        lib.spectral_bwd(w["x"], rw, rmu, rv, w["dk"], w["ec"], k.raw_mixture_weights.grad, k.raw_mixture_means.grad.view(self.Q, -1),
                         k.raw_mixture_scales.grad.view(self.Q, -1), dz, 1, N, self.D, self.Q, self.D, 1, st)
        self.mean_module.constant.grad.copy_(w["hyper"][0, 1:2])
        self.likelihood.noise_covar.raw_noise.grad.copy_(w["hyper"][0, 2:3])
        x4, h, z4 = self._saved
        l1, l2 = self.net.layer1, self.net.layer2
        scratch = torch.empty(lib.conv2d_wgrad_nsplit(N) * 40 * 40, device=x.device)
        gz = dz.view(N, 1, 1, 40)
        lib.conv2d_wgrad(h, gz, z4, l2.weight.grad.view(40, 40, 1, 1), l2.bias.grad, scratch, N, 1, 1, 40, 40, 1, 1, 1, 0, 1,
                         1, st)
        gh = torch.empty_like(h)
        lib.conv2d_dgrad(gz, z4, l2.weight.data.view(40, 40, 1, 1), gh, N, 1, 1, 40, 40, 1, 1, 1, 0, 1, 1, st)
        lib.conv2d_wgrad(x4, gh, h, l1.weight.grad.view(40, 1, 1, 1), l1.bias.grad, scratch, N, 1, 1, 1, 40, 1, 1, 1, 0, 1, 1,
                         st)
        ad = self._adam
        ad["step"] += 1
        pk = self._pack
        lib.adam_step(pk.flat, pk.grad, ad["m"], ad["v"], pk.numel, self.lr, 0.9, 0.999, 1e-8, ad["step"], 1.0, st)
        self._last_info = w["info"]
        return w["loss"][0].clone()

    def predict(self, x_support, y_support, x_query):
        """likelihood(gp(z_query)) conditioned on the support set: (mean, variance incl. noise) [M]."""
        self._ensure()
        lib, st = self.lib, _stream(x_support.device)
        z_s = self._features(x_support.float(), keep=False).clone()
        z_q = self._features(x_query.float(), keep=False).clone()
        N, M = z_s.shape[0], z_q.shape[0]
        w = self._fit(z_s, y_support.float(), want_grad=False)
        rw, rmu, rv = self._spectral()
        kx = torch.empty(1, 1, M, N, device=z_s.device)
        lib.spectral_fwd(z_q.view(1, M, self.D), w["x"], rw, rmu, rv, kx, None, 1, M, N, self.D, self.Q, self.D, 1, st)
        mean = torch.empty(1, 1, M, device=z_s.device)
        var = torch.empty(1, 1, M, device=z_s.device)
        cst = self.mean_module.constant.data.view(1)
        rn = self.likelihood.noise_covar.raw_noise.data.view(1)
        lib.gp_predict(kx, M * N, w["alpha"], None, cst, mean, None, 1, 1, M, N, st)
        kss = torch.nn.functional.softplus(rw).sum().expand(1, 1, M).contiguous()
        lib.gp_predict_var(kx, M * N, kss, M, w["linv"], None, rn, var, 1, 1, M, N, st)
        return mean.view(M), var.view(M)


def main(tot_iterations=50000, n_shot_train=10, n_shot_test=5, test_tasks=500, device="cuda", seed=0):
    """The reference script's train + test protocol (no plots).  Returns the mean test MSE."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    tasks = Task_Distribution()
    model = SinesDKT().to(device)
    for epoch in range(tot_iterations):
        inputs, labels = tasks.sample_task().sample_data(n_shot_train, noise=0.1)
        loss = model.train_step(inputs.to(device), labels.to(device))
        if epoch % 100 == 0:
            print('[%d] - Loss: %.3f  noise: %.3f' % (epoch, loss.item(), model.likelihood.noise.item()))
    mse_list = []
    for _ in range(test_tasks):
        x_all, y_all = tasks.sample_task().sample_data(200, noise=0.1, sort=True)
        idx = np.arange(200)
        np.random.shuffle(idx)
        s, q = np.sort(idx[:n_shot_test]), np.sort(idx[n_shot_test:])
        mean, _ = model.predict(x_all[s].to(device), y_all[s].to(device), x_all[q].to(device))
        mse_list.append(float(((mean.cpu() - y_all[q]) ** 2).mean()))
    print("-------------------")
    print("Average MSE: " + str(np.mean(mse_list)) + " +- " + str(np.std(mse_list)))
    print("-------------------")
    return float(np.mean(mse_list))


if __name__ == "__main__":
    main()
