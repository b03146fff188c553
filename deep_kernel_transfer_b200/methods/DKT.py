"""Deep Kernel Transfer, few-shot classification -- drop-in for the reference ``methods/DKT.py``
(class ``DKT``: __init__ 33-50, get_model_likelihood_mll 58-71, train_loop 113-197, correct 199-272,
test_loop 274-295, get_logits 297-335) with GPyTorch and cuDNN removed from the hot path: every
arithmetic step runs in the hand-written sm_100a kernels behind include/dktb200.h.

Same constructor / methods / attributes / printed fields as the reference, so the reference's
``train.py`` / ``test.py`` / ``test_uncertainty.py`` run against it unmodified.  Extensions (all optional,
defaults reproduce the reference): ``kernel=`` overrides ``configs.kernel_type``; ``episodes_per_step=E``
packs E loader episodes into one meta-step (mean of the per-episode losses; E=1 is the reference's
one-episode-per-Adam-step); under ``torch.distributed`` the flat gradient buffer is all-reduced once
per step (episode-level data parallelism).
"""
import numpy as np
import torch
import torch.nn as nn
from time import gmtime, strftime

from .. import _lib, configs, gp_modules as gpm
from ..engine import GPHead, GPHeadParams, ConvNetParams, check_info, make_targets, _stream
from ..flat import FlatPack
from .meta_template import MetaTemplate

kernel_type = configs.kernel_type      # read at import time, like the reference (DKT.py:14)

try:
    from tensorboardX import SummaryWriter
    IS_TBX_INSTALLED = True
except ImportError:
    IS_TBX_INSTALLED = False


class _Range:
    """NVTX range around one phase of the step (shows up in nsys / ncu --nvtx; a few hundred ns when no tool listens)."""

    def __init__(self, name, on):
        self.name, self.on = name, on

    def __enter__(self):
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


class _DevicePacks:
    """Adapter giving a generator of device-resident packs the prefetcher's interface."""

    def __init__(self, gen):
        self.gen = gen

    def __iter__(self):
        return self.gen

    def release(self, buf):
        pass


class DKT(MetaTemplate):
    def __init__(self, model_func, n_way, n_support, kernel=None, episodes_per_step=1, lib=None):
        super(DKT, self).__init__(model_func, n_way, n_support)
        self.kernel = kernel if kernel is not None else kernel_type
        if self.kernel == "RBF":
            self.kernel = "rbf"
        self.iteration = 0
        self.writer = None
        self.feature_extractor = self.feature
        self.get_model_likelihood_mll()
        if self.kernel == "cossim":
            self.normalize = True
        elif self.kernel == "bncossim":
            self.normalize = True
            latent_size = int(np.prod(self.feature_extractor.final_feat_dim))
            self.feature_extractor.trunk.add_module("bn_out", nn.BatchNorm1d(latent_size))
        else:
            self.normalize = False
        if self.kernel not in ("cossim", "bncossim", "linear", "rbf", "matern", "poli1", "poli2"):
            raise ValueError("[ERROR] the kernel '" + str(self.kernel) + "' is not supported!")
        self.episodes_per_step = int(episodes_per_step)
        self.monitor = True
        self._lib = lib
        self._pack = None
        self._head = None
        self._adam = None
        self.last_step = None

    # ------------------------------------------------------------------ construction (DKT.py:52-71)
    def init_summary(self):
        if IS_TBX_INSTALLED:
            time_string = strftime("%d%m%Y_%H%M%S", gmtime())
            self.writer = SummaryWriter(log_dir="./log/" + time_string)

    def get_model_likelihood_mll(self, train_x_list=None, train_y_list=None):
        models, likelihoods = [], []
        for _ in range(self.n_way):
            likelihood = gpm.GaussianLikelihood(noise=0.1, learn=False)   # DKT.py:346-347
            models.append(gpm.ExactGPLayer(likelihood, kernel=self.kernel))
            likelihoods.append(likelihood)
        self.model = gpm.IndependentModelList(*models)
        self.likelihood = gpm.LikelihoodList(*likelihoods)
        self.mll = gpm.SumMarginalLogLikelihood(self.likelihood, self.model)
        return self.model, self.likelihood, self.mll

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts checkpoints written by the reference (GPyTorch containers, train.py:57-65) as well as this
        package's own: see gp_modules.adapt_reference_state.  Loading copies INTO the flat-buffer views."""
        return super().load_state_dict(gpm.adapt_reference_state(state_dict, self.state_dict()), strict=strict, **kw)

    # ------------------------------------------------------------------ flat parameter buffers
    def _device(self):
        return next(self.feature.parameters()).device

    def _ensure_packed(self):
        dev = self._device()
        if self._pack is not None and self._pack.intact() and self._bufs.intact():
            return
        if dev.type != "cuda" and self._lib is None:
            raise RuntimeError("dktb200 has no CPU path: call .cuda() on the model first")
        lib = self._lib or _lib.load()
        self.lib = lib
        C = self.n_way
        ms = self.model.models
        train = [("gp.raw_outputscale.%d" % c, ms[c].covar_module.raw_outputscale, 1) for c in range(C)]
        train += [("gp.constant.%d" % c, ms[c].mean_module.constant, 1) for c in range(C)]
        self._family = self.kernel not in ("cossim", "bncossim")
        if self._family:      # per-class raw variance / lengthscale / offset (learned, DKT.py:352-365)
            pname = {"linear": "raw_variance", "rbf": "raw_lengthscale", "matern": "raw_lengthscale",
                     "poli1": "raw_offset", "poli2": "raw_offset"}[self.kernel]
            train += [("gp.raw_param.%d" % c, getattr(ms[c].covar_module.base_kernel, pname), 1) for c in range(C)]
        self._gp_count = len(train)
        # every backbone parameter (named_parameters de-duplicates the ConvBlock / feature_extractor aliases; bn_out is
        # part of the trunk) and every floating-point buffer (BatchNorm running statistics)
        train += [("bb." + n, p_, 4) for n, p_ in self.feature.named_parameters()]
        bn_out = getattr(self.feature.trunk, "bn_out", None)
        bufs = [("gp.raw_noise.%d" % c, ms[c].likelihood.noise_covar.raw_noise, 1) for c in range(C)]
        bufs += [("bb." + n, b_, 4) for n, b_ in self.feature.named_buffers() if b_.is_floating_point()]
        # ONE communication buffer [gradients | BatchNorm running statistics + fixed noise]: under torch.distributed a
        # single all-reduce per meta-step covers both (the collective is latency-bound at 0.47 MB)
        n_train, n_bufs = FlatPack.size_of(train), FlatPack.size_of(bufs)
        self._comm = torch.zeros(n_train + n_bufs, device=dev, dtype=torch.float32)
        self._pack = FlatPack(train, dev, grad_storage=self._comm[:n_train])
        self._bufs = FlatPack(bufs, dev, with_grad=False, flat_storage=self._comm[n_train:])
        gp_end = self._pack.offsets[self._gp_count]
        self._gp_range = (0, gp_end)
        self._bb_range = (gp_end, self._pack.numel)
        pk = self._pack
        # engine-facing parameter / gradient views (the modules' storage now IS the flat buffer)
        self._convnet = hasattr(self.feature, "engine_params")
        if self._convnet:
            blocks = self.feature.blocks()
            P, G = ConvNetParams(len(blocks)), ConvNetParams(len(blocks))
            for i, b in enumerate(blocks):
                P.conv_w[i], P.conv_b[i], P.bn_w[i], P.bn_b[i] = b.C.weight.data, b.C.bias.data, b.BN.weight.data, b.BN.bias.data
                G.conv_w[i], G.conv_b[i], G.bn_w[i], G.bn_b[i] = b.C.weight.grad, b.C.bias.grad, b.BN.weight.grad, b.BN.bias.grad
                P.bn_rm[i], P.bn_rv[i] = b.BN.running_mean, b.BN.running_var
            self._P, self._G = P, G
        HP, GH = GPHeadParams(), GPHeadParams()
        HP.raw_outputscale = pk.flat[0:C]
        HP.constant = pk.flat[C:2 * C]
        GH.raw_outputscale = pk.grad[0:C]
        GH.constant = pk.grad[C:2 * C]
        HP.raw_noise = self._bufs.flat[0:C]
        if self._family:
            HP.raw_param = pk.flat[2 * C:3 * C]
            GH.raw_param = pk.grad[2 * C:3 * C]
        if bn_out is not None:
            HP.bn_w, HP.bn_b = bn_out.weight.data, bn_out.bias.data
            GH.bn_w, GH.bn_b = bn_out.weight.grad, bn_out.bias.grad
            HP.bn_rm, HP.bn_rv = bn_out.running_mean, bn_out.running_var
        self._HP, self._GH = HP, GH
        self._head = None
        self._adam = None
        self._graphs = {}            # captured steps hold the old buffers' addresses
        self._sync_replicas()

    def _sync_replicas(self):
        """Data parallelism averages GRADIENTS, so every rank must start from rank 0's parameters / buffers: the
        reference's default ``--seed 0`` means "do not seed" (train.py), i.e. each torchrun rank would otherwise build
        a different random backbone and the replicas would silently diverge."""
        dist, world = self._world()
        if world > 1:
            dist.broadcast(self._pack.flat, 0)
            dist.broadcast(self._bufs.flat, 0)

    def _bb_forward(self, x_all, ipe, training):
        """Backbone forward over packed episodes -> (engine, features [B, D])."""
        eng = self.feature.engine(x_all.shape[-1], x_all.device, self.lib)
        if self._convnet:
            return eng, eng.forward(x_all, self._P, ipe=ipe, training=training)
        feats = eng.forward(x_all, ipe=ipe, training=training)
        if not training:
            eng.tape = []
        return eng, feats

    def _bb_backward(self, eng, x_all, gfeat, ipe):
        if self._convnet:
            eng.backward(x_all, gfeat, self._P, self._G, ipe=ipe)
        else:
            eng.backward(gfeat)

    def _get_head(self, eng):
        if self._head is None or self._head.dev != eng.dev or self._head.D != eng.D:
            self._head = GPHead(self.lib, self.kernel, self.n_way, eng.D, 64 if eng.P > 1 else eng.D, eng.P, eng.dev)
        return self._head

    def _new_adam(self):
        """A fresh Adam at every train_loop call: GP lr 1e-4, backbone lr 1e-3 (DKT.py:114-115)."""
        n = self._pack.numel
        dev = self._pack.flat.device
        ad = self._adam
        if ad is not None and ad["m"].numel() == n and ad["m"].device == dev:
            # in place: a captured CUDA graph of the step holds these addresses
            ad["m"].zero_(); ad["v"].zero_(); ad["step_dev"].zero_()
            ad["step"] = 0
            return
        self._adam = {"m": torch.zeros(n, device=dev), "v": torch.zeros(n, device=dev), "step": 0,
                      "step_dev": torch.zeros(1, device=dev, dtype=torch.int32), "lr_gp": 1e-4, "lr_bb": 1e-3}

    # ------------------------------------------------------------------ one packed meta-train step
    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size()
        return None, 1

    cuda_graph = False      # opt-in: replay the whole meta-step from a CUDA graph (fixed shapes, single process)

    def train_step(self, x_dev):
        """x_dev [E, C, S+Q, 3, H, W] on the device.  Steps 1-6 of the reference's loop body
        (DKT.py:117-193) for E packed episodes.  Returns device tensors (no host sync).

        With ``cuda_graph = True`` the third call with a given shape captures the step (~110 kernel launches + the
        glue ops) into a CUDA graph and later calls replay it: one launch per step instead of ~110 ctypes calls, which is
        what bounds the reference's own granularity of one episode per step (E = 1)."""
        if not (self.cuda_graph and x_dev.is_cuda and self._world()[1] == 1):
            return self._train_step_eager(x_dev)
        self._ensure_packed()
        if self._adam is None:
            self._new_adam()
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (tuple(x_dev.shape), x_dev.device.index, bool(self.monitor))
        st = graphs.setdefault(key, {"calls": 0})
        st["calls"] += 1
        if st["calls"] <= 2:      # eager first (real steps): workspaces, kernel attributes, lazy module loading
            return self._train_step_eager(x_dev)
        if "graph" not in st:
            st["x"] = torch.empty_like(x_dev)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(x_dev.device)
            l0, step0 = self.lib.launches, self._adam["step"]
            with torch.cuda.graph(g):
                st["out"] = self._train_step_eager(st["x"])
            st["graph"], st["launches"] = g, self.lib.launches - l0
            self.lib.launches = l0
            self._adam["step"] = step0            # capturing records, it does not execute
        st["x"].copy_(x_dev)
        st["graph"].replay()
        self.lib.launches += st["launches"]
        self._adam["step"] += 1
        self.last_step = st["out"]
        return st["out"]

    def _train_step_eager(self, x_dev):
        self._ensure_packed()
        if self._adam is None:
            self._new_adam()
        lib = self.lib
        E, C, SQ = x_dev.shape[0], x_dev.shape[1], x_dev.shape[2]
        N = C * SQ
        dev = x_dev.device
        st = _stream(dev)
        x_all = x_dev.reshape(E * N, *x_dev.shape[3:])
        dist, world = self._world()
        targets = self._targets(C, SQ, dev)
        nv = dev.type == "cuda"
        # 1-3: train-mode features, prior, -mll
        with _Range("dkt.backbone_fwd", nv):
            eng, feats = self._bb_forward(x_all, N, True)
        head = self._get_head(eng)
        head.ensure(E, N)
        with _Range("dkt.gp_fit", nv):
            zh = head.embed(feats, self._HP, E, N, training=True, out=head.w["zh_train"])
            loss = head.fit(zh, targets, self._HP, E, N, want_grad=True, grad_scale=1.0 / E)
        # 4: backward + Adam
        with _Range("dkt.gp_bwd", nv):
            gfeat = head.backward(feats, zh, self._HP, self._GH, E, N)
        with _Range("dkt.backbone_bwd", nv):
            self._bb_backward(eng, x_all, gfeat, N)
        if world > 1:      # gradients (summed; Adam divides by world) and running statistics (averaged below) in one collective
            with _Range("dkt.allreduce", nv):
                dist.all_reduce(self._comm)
            lib.scale(self._bufs.flat, self._bufs.numel, 1.0 / world, st)
        ad = self._adam
        ad["step"] += 1
        lib.counter_add(ad["step_dev"], 1, st)        # the step count lives on the device (CUDA-graph replay)
        (g0, g1), (b0, b1) = self._gp_range, self._bb_range
        fl, gr = self._pack.flat, self._pack.grad
        if g1 > g0:
            lib.adam_step_dev(fl[g0:g1], gr[g0:g1], ad["m"][g0:g1], ad["v"][g0:g1], g1 - g0, ad["lr_gp"], 0.9, 0.999, 1e-8,
                              ad["step_dev"], 1.0 / world, st)
        lib.adam_step_dev(fl[b0:b1], gr[b0:b1], ad["m"][b0:b1], ad["v"][b0:b1], b1 - b0, ad["lr_bb"], 0.9, 0.999, 1e-8,
                          ad["step_dev"], 1.0 / world, st)
        for m_ in self.feature.modules():
            if isinstance(m_, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m_.num_batches_tracked += E
        out = {"loss": loss.clone(), "info": head.w["info"]}
        if self.monitor:
            with _Range("dkt.monitor", nv):
                out.update(self.monitor_step(x_dev))
        self.last_step = out
        return out

    def _targets(self, C, per_class, dev):
        """+-1 target matrix [C, C*per_class] on the device, built once per shape (not per step)."""
        key = ("t", C, per_class, str(dev))
        cache = self.__dict__.setdefault("_const_cache", {})
        if key not in cache:
            cache[key] = make_targets(C, per_class, dev)
        return cache[key]

    def _labels(self, C, per_class, dev):
        key = ("l", C, per_class, str(dev))
        cache = self.__dict__.setdefault("_const_cache", {})
        if key not in cache:
            cache[key] = torch.arange(C, device=dev, dtype=torch.int32).repeat_interleave(per_class)
        return cache[key]

    def monitor_step(self, x_dev):
        """Steps 5-6 of the loop body (DKT.py:170-193): eval-mode features of the same images; the GP stays
        conditioned on the pre-update train-mode features kept by ``train_step`` but uses the post-update
        hyper-parameters (Appendix B.4).  Returns per-episode support / query accuracies and the means."""
        E, C, SQ = x_dev.shape[0], x_dev.shape[1], x_dev.shape[2]
        N = C * SQ
        dev = x_dev.device
        x_all = x_dev.reshape(E * N, *x_dev.shape[3:])
        targets = self._targets(C, SQ, dev)
        eng, feats_e = self._bb_forward(x_all, N, False)
        head = self._get_head(eng)
        zh_e = head.embed(feats_e, self._HP, E, N, training=False, out=head.w["zh"])
        head.fit(head.w["zh_train"], targets, self._HP, E, N, want_grad=False)
        if "mon_mean" not in head.w or head.w["mon_mean"].shape != (E, C, N):
            head.w["mon_mean"] = torch.empty(E, C, N, device=dev)
            head.w["mon_pred"] = torch.empty(E, N, device=dev, dtype=torch.int32)
            head.w["mon_kx"] = torch.empty(E, N, N, device=dev)
        head.predict(zh_e, head.w["zh_train"], self._HP, E, N, N, head.w["mon_mean"], head.w["mon_pred"],
                     head.w["mon_kx"])
        labels = self._labels(C, SQ, dev)
        hit = (head.w["mon_pred"] == labels.unsqueeze(0)).view(E, C, SQ)
        return {"acc_support": hit[:, :, :self.n_support].float().mean((1, 2)) * 100.0,
                "acc_query": hit[:, :, self.n_support:].float().mean((1, 2)) * 100.0,
                "hit_support": hit[:, :, :self.n_support].sum((1, 2)), "hit_query": hit[:, :, self.n_support:].sum((1, 2)),
                "mean": head.w["mon_mean"], "pred": head.w["mon_pred"]}

    # ------------------------------------------------------------------ reference API
    def train_loop(self, epoch, train_loader, optimizer=None, print_freq=10):
        # the reference ignores `optimizer` and builds its own Adam here, every call (DKT.py:114-115)
        self._ensure_packed()
        self._new_adam()
        dev = self._device()
        n_items = len(train_loader)
        E = self.episodes_per_step

        def packs():
            pend = []
            for i, (x, _) in enumerate(train_loader):
                pend.append(x)
                if len(pend) < E and i + 1 < n_items:
                    continue
                xs = torch.stack(pend, 0) if len(pend) > 1 else pend[0].unsqueeze(0)
                pend = []
                yield xs

        from ..feeder import DevicePrefetcher
        if hasattr(train_loader, "device_packs"):   # GPU episode feeder: packs are assembled on the device (episode_feed.py)
            feed = _DevicePacks(train_loader.device_packs(E))
        else:
            feed = DevicePrefetcher(packs(), dev)   # H2D of pack k+1 overlaps the kernels of pack k
        step_i = 0
        pending = []            # (iteration, loss, hit_support, hit_query, n_s, n_q): device scalars, logged in order
        self._sync_replicas()

        def flush():
            """The reference writes loss and both accuracies to the summary writer at EVERY iteration
            (DKT.py:167, 183, 193).  Same records, same order, same iteration numbers -- delivered in batches (at the
            print steps and at the end of the loop) so that logging does not force a host sync per step."""
            if self.writer is not None:
                for it, loss_t, hs, hq, n_s, n_q in pending:
                    self.writer.add_scalar("loss", float(loss_t), it)
                    if hs is not None:
                        self.writer.add_scalar("GP_support_accuracy", (float(hs) / n_s) * 100.0, it)
                        self.writer.add_scalar("GP_query_accuracy", (float(hq) / n_q) * 100.0, it)
            del pending[:]

        def health():
            """Cholesky status of every fit since the last look + the tcgen05 pipeline time-out flag: one read-back."""
            if self._head is not None:
                self._head.check()
            eng = getattr(self.feature, "_engine", None)
            if eng is not None and hasattr(eng, "check_tc"):
                eng.check_tc()

        for x_dev in feed:
            i = min((step_i + 1) * E, n_items) - 1
            self.n_query = x_dev.size(2) - self.n_support
            if self.change_way:
                self.n_way = x_dev.size(1)
            out = self.train_step(x_dev)
            feed.release(x_dev)
            self.iteration = i + (epoch * n_items)
            n_s = float(x_dev.size(0) * x_dev.size(1) * self.n_support)
            n_q = float(x_dev.size(0) * x_dev.size(1) * self.n_query)
            if self.writer is not None:
                pending.append((self.iteration, out["loss"].mean(),
                                out["hit_support"].sum() if self.monitor else None,
                                out["hit_query"].sum() if self.monitor else None, n_s, n_q))
            if step_i % print_freq == 0:
                health()
                ms = self.model.models
                outputscale = float(np.mean([m.covar_module.outputscale.item() for m in ms]))
                noise = float(np.mean([m.likelihood.noise.item() for m in ms]))
                ls = [m.covar_module.base_kernel.lengthscale for m in ms]
                lenghtscale = float(np.mean([l.mean().item() for l in ls])) if ls[0] is not None else 0.0
                loss = out["loss"].mean().item()
                # percentages from the integer hit counts in double, as the reference's numpy does (DKT.py:181, 191)
                acc_s = (out["hit_support"].sum().item() / n_s) * 100.0 if self.monitor else float("nan")
                acc_q = (out["hit_query"].sum().item() / n_q) * 100.0 if self.monitor else float("nan")
                flush()
                print('Epoch [{:d}] [{:d}/{:d}] | Outscale {:f} | Lenghtscale {:f} | Noise {:f} | Loss {:f} | Supp. {:f} | Query {:f}'.format(
                    epoch, i, n_items, outputscale, lenghtscale, noise, loss, acc_s, acc_q))
            step_i += 1
        flush()
        health()

    def _episode_embed(self, x):
        """x [C, S+Q, 3, H, W] (any device) -> (engine, support embedding [1, C*S, D], query embedding [1, C*Q, D]):
        eval-mode backbone (+ bn_out, + F.normalize for the cosine kernels), DKT.py:236-237 / 262-263."""
        self._ensure_packed()
        dev = self._device()
        C, SQ = x.shape[0], x.shape[1]
        S = self.n_support
        x_dev = x.to(dev, non_blocking=True).float().contiguous().view(C * SQ, *x.shape[2:])
        B = C * SQ
        eng, feats = self._bb_forward(x_dev, B, False)                      # eval-mode BN is per-sample
        head = self._get_head(eng)
        head.ensure(1, B)
        zh_all = head.embed(feats, self._HP, 1, B, training=False, out=head.w["zh"])[0]       # [B, D]
        idx = torch.arange(B, device=dev).view(C, SQ)
        zh_s = zh_all.index_select(0, idx[:, :S].reshape(-1)).unsqueeze(0).contiguous()       # [1, C*S, D]
        zh_q = zh_all.index_select(0, idx[:, S:].reshape(-1)).unsqueeze(0).contiguous()       # [1, C*Q, D]
        return eng, zh_s, zh_q

    def _episode_logits(self, x, adapt_steps=0):
        """x [C, S+Q, 3, H, W] (any device) -> (mean [C, M] device tensor, pred [M] int32 device tensor).
        adapt_steps > 0: that many Adam(lr 1e-3) steps on the GP hyper-parameters first (DKT.py:241-256)."""
        eng, zh_s, zh_q = self._episode_embed(x)
        dev = zh_s.device
        C, S = x.shape[0], self.n_support
        Q = x.shape[1] - S
        N, M = C * S, C * Q
        fit_head = self._test_head(eng, N)
        self._adapt_loss = 0.0
        if adapt_steps > 0:
            # test-time adaptation of the GP hyper-parameters on the (eval-mode, constant) support features; a fresh
            # Adam over the GP slice of the flat buffer, and -- as in the reference -- the updated values stay in the model
            g0, g1 = self._gp_range
            st = _stream(dev)
            m_, v_ = torch.zeros(g1 - g0, device=dev), torch.zeros(g1 - g0, device=dev)
            tg = make_targets(C, S, dev)
            total = torch.zeros((), device=dev)
            for it in range(adapt_steps):
                loss = fit_head.fit(zh_s, tg, self._HP, 1, N, want_grad=True, grad_scale=1.0)
                fit_head.hyper_grads(self._HP, self._GH, 1, N)
                total += loss[0]
                self.lib.adam_step(self._pack.flat[g0:g1], self._pack.grad[g0:g1], m_, v_, g1 - g0, 1e-3, 0.9, 0.999, 1e-8,
                                   it + 1, 1.0, st)
            check_info(fit_head.w["info"])
            self._adapt_loss = float(total) / float(adapt_steps + 1e-10)
        fit_head.fit(zh_s, make_targets(C, S, dev), self._HP, 1, N, want_grad=False)
        mean = torch.empty(1, C, M, device=dev)
        pred = torch.empty(1, M, device=dev, dtype=torch.int32)
        kx = torch.empty(1, M, N, device=dev)
        fit_head.predict(zh_q, zh_s, self._HP, 1, M, N, mean, pred, kx)
        self._last_info = fit_head.w["info"]
        self._last_mean, self._last_pred = mean[0], pred[0]
        return eng, pred[0]

    def _test_head(self, eng, N):
        h = getattr(self, "_thead", None)
        if h is None or h.dev != eng.dev or h.D != eng.D:
            h = GPHead(self.lib, self.kernel, self.n_way, eng.D, 64 if eng.P > 1 else eng.D, eng.P, eng.dev)
            self._thead = h
        h.ensure(1, N)
        return h

    def correct(self, x, N=0, laplace=False):
        C = x.shape[0]
        self.n_query = x.size(1) - self.n_support
        if laplace:
            # DKT.py:207-224: the reference itself leaves the GP here -- the (CUDA) embeddings go to scikit-learn's
            # GaussianProcessClassifier (Laplace approximation, fixed 1.0 * RBF(0.1) kernel, no optimisation) on the host
            from sklearn.gaussian_process import GaussianProcessClassifier
            from sklearn.gaussian_process.kernels import RBF
            _, zh_s, zh_q = self._episode_embed(x)
            gp = GaussianProcessClassifier(kernel=1.0 * RBF(length_scale=0.1, length_scale_bounds=(0.1, 10.0)), optimizer=None)
            gp.fit(zh_s[0].cpu().numpy(), np.repeat(range(C), self.n_support))
            y_pred = gp.predict(zh_q[0].cpu().numpy())
            y_query = np.repeat(range(C), self.n_query)
            return float(np.sum(y_pred == y_query)), len(y_query), 0.0
        eng, pred = self._episode_logits(x, adapt_steps=int(N))
        check_info(self._last_info)
        if hasattr(eng, "check_tc"):
            eng.check_tc()
        y_query = np.repeat(range(C), self.n_query)
        top1_correct = np.sum(pred.cpu().numpy() == y_query)
        return float(top1_correct), len(y_query), self._adapt_loss

    def correct_packed(self, xs):
        """E test episodes in ONE kernel chain: xs [E, C, S+Q, 3, H, W] (any device) -> int64 tensor [E] of top-1 hits
        (host).  Same arithmetic per episode as ``correct(x)`` -- eval-mode BatchNorm is per sample and every episode has
        its own C exact-GP systems -- so the counts are identical; what changes is one launch chain and one host sync
        per E episodes instead of per episode (600 x 5 test episodes in test.py:65 are launch-latency-bound otherwise)."""
        self._ensure_packed()
        dev = self._device()
        E, C, SQ = xs.shape[0], xs.shape[1], xs.shape[2]
        S = self.n_support
        Q = SQ - S
        B, N, M = C * SQ, C * S, C * Q
        x_dev = xs.to(dev, non_blocking=True).float().contiguous().view(E * B, *xs.shape[3:])
        eng, feats = self._bb_forward(x_dev, B, False)
        head = self._packed_test_head(eng, E, B)
        zh_all = head.embed(feats, self._HP, E, B, training=False, out=head.w["zh"])            # [E, B, D]
        idx = torch.arange(B, device=dev).view(C, SQ)
        zh_s = zh_all.index_select(1, idx[:, :S].reshape(-1)).contiguous()                       # [E, C*S, D]
        zh_q = zh_all.index_select(1, idx[:, S:].reshape(-1)).contiguous()                       # [E, C*Q, D]
        fit_head = self._packed_fit_head(eng, E, N)
        fit_head.fit(zh_s, self._targets(C, S, dev), self._HP, E, N, want_grad=False)
        t = fit_head.w.setdefault("packed_out", {})
        if t.get("key") != (E, M, N):
            t.update(key=(E, M, N), mean=torch.empty(E, C, M, device=dev), pred=torch.empty(E, M, device=dev, dtype=torch.int32),
                     kx=torch.empty(E, M, N, device=dev))
        fit_head.predict(zh_q, zh_s, self._HP, E, M, N, t["mean"], t["pred"], t["kx"])
        hits = (t["pred"] == self._labels(C, Q, dev).unsqueeze(0)).sum(1)
        fit_head.check()
        if hasattr(eng, "check_tc"):
            eng.check_tc()
        self._last_mean, self._last_pred = t["mean"], t["pred"]
        return hits.cpu()

    def _packed_test_head(self, eng, E, B):
        h = getattr(self, "_phead", None)
        if h is None or h.dev != eng.dev or h.D != eng.D:
            h = GPHead(self.lib, self.kernel, self.n_way, eng.D, 64 if eng.P > 1 else eng.D, eng.P, eng.dev)
            self._phead = h
        h.ensure(E, B)
        return h

    def _packed_fit_head(self, eng, E, N):
        h = getattr(self, "_pfit", None)
        if h is None or h.dev != eng.dev or h.D != eng.D:
            h = GPHead(self.lib, self.kernel, self.n_way, eng.D, 64 if eng.P > 1 else eng.D, eng.P, eng.dev)
            self._pfit = h
        h.ensure(E, N)
        return h

    test_episodes_per_call = 25      # packing factor of test_loop (results do not depend on it)

    def test_loop(self, test_loader, record=None, return_std=False):
        acc_all = []
        iter_num = len(test_loader)
        E = max(1, int(self.test_episodes_per_call))
        pend = []

        def run(batch):
            # stack into a pinned staging buffer (one H2D per pack at the host link's rate instead of a pageable copy);
            # correct_packed ends with a host sync, so the buffer is free again when the next pack is staged
            shape = (E,) + tuple(batch[0].shape)
            stage = self.__dict__.get("_test_stage")
            if stage is None or tuple(stage.shape) != shape:
                stage = torch.empty(shape, dtype=torch.float32)
                if self._device().type == "cuda":
                    stage = stage.pin_memory()
                self._test_stage = stage
            xs = stage[:len(batch)]
            torch.stack([b.float() for b in batch], 0, out=xs)
            hits = self.correct_packed(xs)
            count_this = xs.size(1) * (xs.size(2) - self.n_support)
            return [float(h) / count_this * 100 for h in hits.tolist()]

        def emit(accs, first_index):
            for j, a in enumerate(accs):
                acc_all.append(a)
                i = first_index + j
                if i % 100 == 0:
                    acc_mean = np.mean(np.asarray(acc_all))
                    print('Test | Batch {:d}/{:d} | Loss {:f} | Acc {:f}'.format(i, len(test_loader), 0.0, acc_mean))
        first = 0
        for i, (x, _) in enumerate(test_loader):
            self.n_query = x.size(1) - self.n_support
            if self.change_way:
                self.n_way = x.size(0)
            if pend and (tuple(x.shape) != tuple(pend[0].shape) or len(pend) == E):
                emit(run(pend), first)
                first, pend = i, []
            pend.append(x)
        if pend:
            emit(run(pend), first)
        acc_all = np.asarray(acc_all)
        acc_mean = np.mean(acc_all)
        acc_std = np.std(acc_all)
        print('%d Test Acc = %4.2f%% +- %4.2f%%' % (iter_num, acc_mean, 1.96 * acc_std / np.sqrt(iter_num)))
        if self.writer is not None:
            self.writer.add_scalar('test_accuracy', acc_mean, self.iteration)
        if return_std:
            return acc_mean, acc_std
        return acc_mean

    def get_logits(self, x):
        self.n_query = x.size(1) - self.n_support
        self._episode_logits(x)
        return self._last_mean.t().contiguous()          # [C*Q, C]  (torch.stack(means, 1), DKT.py:333-335)
