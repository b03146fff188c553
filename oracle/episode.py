"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of one DKT episode.

Follows the reference control flow:
  * meta-train step  methods/DKT.py:113-197  (steps 1-6 of SURVEY.md 3.1)
  * test episode     methods/DKT.py:199-272  (``correct``) and 297-335 (``get_logits``)
  * regression       methods/DKT_regression.py:45-64 (train), 66-97 (test)
using oracle.backbone (pinned to the reference's backbone.py) and oracle.gp.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import backbone as obb
from . import gp as ogp

NORMALIZED_KERNELS = ("cossim", "bncossim")   # DKT.py:43-50


def make_targets(n_way, per_class, dtype=torch.float32):
    """DKT.py:129-136 / 227-234: one +-1 target vector per class, class-major sample order."""
    t = -torch.ones(n_way, n_way * per_class, dtype=dtype)
    for c in range(n_way):
        t[c, c * per_class:(c + 1) * per_class] = 1.0
    return t


def features(arch, bb, x, kernel, training, update_running=True):
    z = obb.forward(arch, bb, x, training=training, update_running=update_running)
    if kernel in NORMALIZED_KERNELS:
        z = F.normalize(z, p=2, dim=1)        # DKT.py:142 (eps 1e-12)
    return z


class OracleDKT:
    """Holds backbone + GP parameters and an Adam pair exactly like DKT.train_loop (DKT.py:114-115)."""

    def __init__(self, arch="Conv4", kernel="bncossim", n_way=5, n_support=5, seed=0, dtype=torch.float32,
                 feat_dim=None):
        self.arch, self.kernel, self.n_way, self.n_support = arch, kernel, n_way, n_support
        self.bb = obb.init_params(arch, seed=seed, bn_out=(kernel == "bncossim"), bn_out_dim=feat_dim)
        self.bb = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in self.bb.items()}
        self.gp = ogp.default_gp_params(kernel, n_way, feat_dim or obb.feat_dim(arch), dtype=dtype, classification=True)
        self.optimizer = None

    # -- parameter bookkeeping
    def bb_trainable(self):
        return [k for k in self.bb if k.endswith(".weight") or k.endswith(".bias")]

    def new_optimizer(self):
        """A fresh Adam at every train_loop() call: GP lr 1e-4, backbone lr 1e-3 (DKT.py:114-115)."""
        for k in self.bb_trainable():
            self.bb[k].requires_grad_(True)
        for k in ogp.trainable_gp_names(self.kernel):
            self.gp[k].requires_grad_(True)
        self.optimizer = torch.optim.Adam(
            [{"params": [self.gp[k] for k in ogp.trainable_gp_names(self.kernel)], "lr": 1e-4},
             {"params": [self.bb[k] for k in self.bb_trainable()], "lr": 1e-3}])

    # -- one meta-train step over a pack of E episodes (E=1 is the reference's behaviour)
    def train_step(self, xs, monitor=True):
        """xs: tensor [E, C, S+Q, 3, H, W] (or [C, S+Q, 3, H, W]).  Returns a dict of results."""
        if xs.dim() == 5:
            xs = xs.unsqueeze(0)
        if self.optimizer is None:
            self.new_optimizer()
        e_count, c, sq = xs.shape[0], xs.shape[1], xs.shape[2]
        n_query = sq - self.n_support
        targets = make_targets(c, sq, xs.dtype)
        self.optimizer.zero_grad()
        losses, z_trains = [], []
        for e in range(e_count):
            x_all = xs[e].reshape(c * sq, *xs.shape[3:])
            z = features(self.arch, self.bb, x_all, self.kernel, training=True)      # step 1
            loss = ogp.mll_loss(self.kernel, z, targets, self.gp)                       # steps 2-3
            (loss / e_count).backward()                                                  # step 4
            losses.append(loss.detach())
            z_trains.append(z.detach())
        grads = {k: self.bb[k].grad.clone() for k in self.bb_trainable()}
        grads.update({k: self.gp[k].grad.clone() for k in ogp.trainable_gp_names(self.kernel)})
        self.optimizer.step()
        out = {"loss": torch.stack(losses), "grads": grads, "z_train": torch.stack(z_trains)}
        if monitor:                                                                      # steps 5-6
            y_s = np.repeat(range(c), self.n_support)
            y_q = np.repeat(range(c), n_query)
            acc_s, acc_q, mean_s, mean_q = [], [], [], []
            with torch.no_grad():
                for e in range(e_count):
                    x_s = xs[e][:, :self.n_support].reshape(c * self.n_support, *xs.shape[3:])
                    x_q = xs[e][:, self.n_support:].reshape(c * n_query, *xs.shape[3:])
                    z_s = features(self.arch, self.bb, x_s, self.kernel, training=False)
                    z_q = features(self.arch, self.bb, x_q, self.kernel, training=False)
                    # GP is still conditioned on the pre-update, train-mode features (Appendix B.4)
                    m_s = ogp.predict(self.kernel, z_trains[e], targets, z_s, self.gp)
                    m_q = ogp.predict(self.kernel, z_trains[e], targets, z_q, self.gp)
                    p_s = torch.sigmoid(m_s).numpy().argmax(axis=0)
                    p_q = torch.sigmoid(m_q).numpy().argmax(axis=0)
                    acc_s.append(np.sum(p_s == y_s) / float(len(y_s)) * 100.0)
                    acc_q.append(np.sum(p_q == y_q) / float(len(y_q)) * 100.0)
                    mean_s.append(m_s)
                    mean_q.append(m_q)
            out.update(acc_support=np.array(acc_s), acc_query=np.array(acc_q),
                       mean_support=torch.stack(mean_s), mean_query=torch.stack(mean_q))
        return out

    # -- test episode (DKT.correct with N=0, laplace=False)
    def get_logits(self, x):
        c, sq = x.shape[0], x.shape[1]
        n_query = sq - self.n_support
        x_s = x[:, :self.n_support].reshape(c * self.n_support, *x.shape[2:])
        x_q = x[:, self.n_support:].reshape(c * n_query, *x.shape[2:])
        targets = make_targets(c, self.n_support, x.dtype)
        with torch.no_grad():
            z_tr = features(self.arch, self.bb, x_s, self.kernel, training=False)
            z_q = features(self.arch, self.bb, x_q, self.kernel, training=False)
            mean = ogp.predict(self.kernel, z_tr, targets, z_q, self.gp)     # [C, M]
        return mean.t().contiguous()                                          # [C*Q, C] (DKT.py:333-335)

    def correct_laplace(self, x):
        """DKT.correct(laplace=True) (methods/DKT.py:207-224): scikit-learn's Laplace-approximated GP classifier with
        the fixed kernel 1.0 * RBF(0.1), no optimisation, on the eval-mode features."""
        from sklearn.gaussian_process import GaussianProcessClassifier
        from sklearn.gaussian_process.kernels import RBF
        c = x.shape[0]
        s = self.n_support
        n_query = x.shape[1] - s
        with torch.no_grad():
            z_s = features(self.arch, self.bb, x[:, :s].reshape(c * s, *x.shape[2:]), self.kernel, training=False)
            z_q = features(self.arch, self.bb, x[:, s:].reshape(c * n_query, *x.shape[2:]), self.kernel, training=False)
        gp = GaussianProcessClassifier(kernel=1.0 * RBF(length_scale=0.1, length_scale_bounds=(0.1, 10.0)), optimizer=None)
        gp.fit(z_s.numpy(), np.repeat(range(c), s))
        y_pred = gp.predict(z_q.numpy())
        y_query = np.repeat(range(c), n_query)
        return float(np.sum(y_pred == y_query)), len(y_query), 0.0

    def correct(self, x, N=0):
        """DKT.correct (methods/DKT.py:199-272, laplace=False).  N > 0: N Adam(lr 1e-3) steps on the GP hyper-parameters
        with the (eval-mode, detached) support features as training data (DKT.py:241-256); the updated hyper-parameters
        stay in the model, as in the reference.  Returns (top1_correct, count, avg_loss)."""
        c = x.shape[0]
        n_query = x.shape[1] - self.n_support
        avg_loss = 0.0
        if N > 0:
            x_s = x[:, :self.n_support].reshape(c * self.n_support, *x.shape[2:])
            targets = make_targets(c, self.n_support, x.dtype)
            with torch.no_grad():
                z_tr = features(self.arch, self.bb, x_s, self.kernel, training=False)
            names = ogp.trainable_gp_names(self.kernel)
            for k in names:
                self.gp[k] = self.gp[k].detach().clone().requires_grad_(True)
            opt = torch.optim.Adam([self.gp[k] for k in names], lr=1e-3)
            for _ in range(N):
                opt.zero_grad()
                loss = ogp.mll_loss(self.kernel, z_tr, targets, self.gp)
                loss.backward()
                opt.step()
                avg_loss += float(loss)
            for k in names:
                self.gp[k] = self.gp[k].detach()
        logits = self.get_logits(x)
        y_pred = torch.sigmoid(logits.t()).numpy().argmax(axis=0)
        y_query = np.repeat(range(c), n_query)
        return float(np.sum(y_pred == y_query)), len(y_query), avg_loss / float(N + 1e-10)


def synthetic_episode(episode_id, n_way=5, n_support=5, n_query=16, image_size=84, dtype=torch.float32):
    """Synthetic episode of SURVEY.md 8(d): x ~ N(0,1) + 0.5 * per-class offset, seed 1234+id."""
    g = torch.Generator().manual_seed(1234 + episode_id)
    x = torch.randn(n_way, n_support + n_query, 3, image_size, image_size, generator=g)
    off = 0.5 * torch.randn(n_way, 1, 3, 1, 1, generator=g)
    return (x + off).to(dtype)


# ----------------------------------------------------------------------------- regression
class OracleDKTRegression:
    """methods/DKT_regression.py: one exact GP (RBF or spectral mixture) on Conv3 features."""

    def __init__(self, kernel="rbf", arch="Conv3", seed=0, dtype=torch.float32, feat_dim=None):
        self.arch, self.kernel = arch, kernel
        self.bb = {k: v.to(dtype) for k, v in obb.init_params(arch, seed=seed).items()}
        d = feat_dim or obb.feat_dim(arch)
        self.gp = ogp.default_gp_params(kernel, 1, d, dtype=dtype, classification=False)
        self.optimizer = None

    def new_optimizer(self, lr_gp=1e-3, lr_net=1e-3):     # train_regression.py:33-34
        for v in self.bb.values():
            v.requires_grad_(True)
        names = ogp.trainable_gp_names(self.kernel, classification=False)
        for k in names:
            self.gp[k].requires_grad_(True)
        self.optimizer = torch.optim.Adam([{"params": [self.gp[k] for k in names], "lr": lr_gp},
                                           {"params": list(self.bb.values()), "lr": lr_net}])

    def train_step(self, inputs, labels):
        """One person: inputs [19,3,100,100], labels [19] (DKT_regression.py:48-57)."""
        if self.optimizer is None:
            self.new_optimizer()
        self.optimizer.zero_grad()
        z = obb.forward(self.arch, self.bb, inputs)
        loss = ogp.mll_loss(self.kernel, z, labels.unsqueeze(0), self.gp)
        loss.backward()
        grads = {k: v.grad.clone() for k, v in self.bb.items()}
        grads.update({k: self.gp[k].grad.clone() for k in ogp.trainable_gp_names(self.kernel, False)})
        self.optimizer.step()
        mse = ((self.gp["constant"][0].detach() - labels) ** 2).mean()   # prior mean (DKT_regression.py:58)
        return {"loss": loss.detach(), "grads": grads, "mse": mse, "z": z.detach()}

    def test_episode(self, x_support, y_support, x_all, y_all):
        """DKT_regression.py:83-95: fit on the support, predict all; returns (mse, mean, lower, upper)."""
        with torch.no_grad():
            z_s = obb.forward(self.arch, self.bb, x_support)
            z_q = obb.forward(self.arch, self.bb, x_all)
            mean, var = ogp.predict(self.kernel, z_s, y_support.unsqueeze(0), z_q, self.gp, want_var=True)
            mean, var = mean[0], var[0]
            std = var.clamp_min(1e-10).sqrt()
            mse = ((mean - y_all) ** 2).mean()
        return mse, mean, mean - 2.0 * std, mean + 2.0 * std


# ----------------------------------------------------------------------------- sines (sines/train_DKT.py)
class OracleSines:
    """MLP 1->40->40 (ReLU) + spectral-mixture GP (Q=4, ARD 40) + learned noise, Adam lr 1e-3 (sines/train_DKT.py:113-180)."""

    def __init__(self, dtype=torch.float32, seed=0):
        g = torch.Generator().manual_seed(seed)
        u = lambda *s, b=1.0: ((torch.rand(*s, generator=g) * 2 - 1) * b).to(dtype)
        self.p = {"layer1.weight": u(40, 1), "layer1.bias": u(40), "layer2.weight": u(40, 40, b=40 ** -0.5),
                  "layer2.bias": u(40, b=40 ** -0.5)}
        self.gp = ogp.default_gp_params("spectral", 1, 40, dtype=dtype, classification=False)
        self.optimizer = None

    def feats(self, x):
        h = F.relu(F.linear(x, self.p["layer1.weight"], self.p["layer1.bias"]))
        return F.relu(F.linear(h, self.p["layer2.weight"], self.p["layer2.bias"]))

    def train_step(self, x, y):
        names = ogp.trainable_gp_names("spectral", classification=False)
        if self.optimizer is None:
            for v in self.p.values():
                v.requires_grad_(True)
            for k in names:
                self.gp[k].requires_grad_(True)
            self.optimizer = torch.optim.Adam([{"params": [self.gp[k] for k in names], "lr": 1e-3},
                                               {"params": list(self.p.values()), "lr": 1e-3}])
        self.optimizer.zero_grad()
        loss = ogp.mll_loss("spectral", self.feats(x), y.unsqueeze(0), self.gp)
        loss.backward()
        grads = {k: v.grad.clone() for k, v in self.p.items()}
        grads.update({k: self.gp[k].grad.clone() for k in names})
        self.optimizer.step()
        return {"loss": loss.detach(), "grads": grads}

    def predict(self, xs, ys, xq):
        with torch.no_grad():
            mean, var = ogp.predict("spectral", self.feats(xs), ys.unsqueeze(0), self.feats(xq), self.gp, want_var=True)
        return mean[0], var[0]
