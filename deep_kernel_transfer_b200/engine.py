"""Host-side orchestration of the dktb200 kernels: packed-episode forward/backward of the ConvNet
backbone (reference backbone.py:105-132, 250-268) and of the exact-GP head (methods/DKT.py:141-164,
170-193, 236-269).  PyTorch is used for device memory, streams and (elsewhere) torch.distributed
only -- every arithmetic step below is a call through the C ABI in include/dktb200.h.

Data layout in HBM (all fp32):
  x        [B,3,H,W]            NCHW exactly as the loader delivers it (B = E episodes x ipe images)
  y[0]     [B,H,W,64]           pre-BN conv1 output (NHWC)
  y[i>0]   [B,H_i+2,W_i+2,64]   pre-BN conv output in the padded-flat layout (border never read)
  act[i]   [B,Ho+2,Wo+2,64]     block output = next conv's input, ZERO border (written once at alloc)
  feats    [B, P*64]            last block output, NHWC-flattened
  gy[i], gact[i]                gradients in the same layouts (gy borders stay zero: dgrad/wgrad read them)
"""
import math

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
GP_SMEM_CROSSOVER = 105   # largest N served by the all-in-shared-memory gp_fit kernel (the tiled one wins from ~100 on)
GP_JITTER = 1e-6          # psd_safe_cholesky retry base for float32 (GPyTorch utils/cholesky.py): 1e-6, 1e-5, 1e-4


def _stream(dev):
    if dev.type == "cuda":
        return torch.cuda.current_stream(dev).cuda_stream
    return 0


class ConvNetParams:
    """Plain container of the tensors one ConvNet needs (views into the flat parameter buffer)."""

    def __init__(self, depth):
        self.depth = depth
        self.conv_w = [None] * depth
        self.conv_b = [None] * depth
        self.bn_w = [None] * depth
        self.bn_b = [None] * depth
        self.bn_rm = [None] * depth
        self.bn_rv = [None] * depth


class ConvNetEngine:
    """Conv4 / Conv6 (``ConvNet(depth)``) over packed episodes."""

    def __init__(self, lib, depth=4, image_size=84, device="cuda", use_tc=None):
        self.lib = lib
        self.depth = depth
        self.dev = torch.device(device)
        # CUDA device: the tcgen05 kernels (csrc/conv_tc.cu, conv1_bwd_mma.cu), always.  The fp32 CUDA-core twins
        # (csrc/conv_fp32.cu) exist for the g++ host build of the kernels that the CPU tests inject: a CPU "device"
        # is the only way to reach them -- there is no switch between code paths on the GPU.
        if use_tc is None:
            use_tc = self.dev.type == "cuda"
        self.use_tc = bool(use_tc)
        if self.use_tc and image_size + 2 > 88:
            raise NotImplementedError("ConvNet inputs wider than 86 pixels (the reference's Conv4 / Conv6 run at 84)")
        self.wgrad_tc = self.use_tc
        self.conv1_tc = self.use_tc
        # first-block backward: BN/ReLU/pool backward fused into the conv1 weight gradient (no gy[0] tensor at all)
        self.l0_fused = True
        self.l0_fn = "conv1_bwd_fused_mma" if self.use_tc else "conv1_bwd_fused"
        self.layers = []
        h = image_size
        for i in range(depth):
            pool = i < 4
            self.layers.append({"H": h, "W": h, "pool": pool})
            if pool:
                h = h // 2
        self.Hf = h
        self.P = h * h
        self.D = 64 * self.P
        self.cap = 0
        self.ws = None

    # ------------------------------------------------------------------ workspace
    def _alloc(self, B, E):
        dev, f32 = self.dev, torch.float32
        ws = {"y": [], "act": [], "gy": [], "gact": [], "mean": [], "invstd": [], "partials": [], "T": [],
              "wt_f": [], "wt_d": []}
        lib = self.lib
        max_part = 0
        for i, L in enumerate(self.layers):
            H, W = L["H"], L["W"]
            Ho, Wo = (H // 2, W // 2) if L["pool"] else (H, W)
            last = i == self.depth - 1
            if i == 0:
                ws["y"].append(torch.empty(B, H, W, 64, device=dev, dtype=f32))
                ws["gy"].append(None if (self.l0_fused and L["pool"]) else torch.empty(B, H, W, 64, device=dev, dtype=f32))
                T = lib.conv1_tiles(H, W)
            else:
                ws["y"].append(torch.zeros(B, H + 2, W + 2, 64, device=dev, dtype=f32))
                ws["gy"].append(torch.zeros(B, H + 2, W + 2, 64, device=dev, dtype=f32))
                T = lib.conv3x3_tiles(H, W)
            if last:
                ws["act"].append(torch.empty(B, Ho, Wo, 64, device=dev, dtype=f32))
                ws["gact"].append(None)     # provided by the head
            else:
                ws["act"].append(torch.zeros(B, Ho + 2, Wo + 2, 64, device=dev, dtype=f32))
                ws["gact"].append(torch.zeros(B, Ho + 2, Wo + 2, 64, device=dev, dtype=f32))
            ws["T"].append(T)
            ws["partials"].append(torch.empty(B * T * 128, device=dev, dtype=f32))
            ws["mean"].append(torch.empty(E, 64, device=dev, dtype=f32))
            ws["invstd"].append(torch.empty(E, 64, device=dev, dtype=f32))
            shape = (lib.conv3x3_tc_weight_floats(),) if self.use_tc else (9, 64, 64)
            ws["wt_f"].append(torch.empty(*shape, device=dev, dtype=f32) if i > 0 else None)
            ws["wt_d"].append(torch.empty(*shape, device=dev, dtype=f32) if i > 0 else None)
            max_part = max(max_part, B * lib.bn_bwd_chunks(H, W, int(L["pool"])) * 128)
        ws["tc_err"] = torch.zeros(1, device=dev, dtype=torch.int32)
        ws["wb1"] = torch.empty(2, 64, 32, device=dev, dtype=f32)
        ws["eval_mean"] = torch.empty(self.depth, 64, device=dev, dtype=f32)
        ws["eval_invstd"] = torch.empty(self.depth, 64, device=dev, dtype=f32)
        ws["bwd_partial"] = torch.empty(max_part, device=dev, dtype=f32)
        ws["bwd_sums"] = torch.empty(E * 128, device=dev, dtype=f32)
        ws["scratch_d"] = torch.empty(lib.bn_scratch_doubles(E), device=dev, dtype=torch.float64)
        n1 = lib.conv1_wgrad_nsplit() * 28 * 64
        ws["wgrad_scratch"] = torch.empty(max(n1, lib.conv3x3_wgrad_scratch_floats()), device=dev, dtype=f32)
        self.ws, self.cap, self.cap_E = ws, B, E

    def ensure(self, B, E):
        if self.ws is None or B != self.cap or E != self.cap_E:
            self._alloc(B, E)

    # ------------------------------------------------------------------ forward
    def prepare_weights(self, P):
        st = _stream(self.dev)
        if self.conv1_tc:
            self.lib.prep_weights_conv1_tc(P.conv_w[0], self.ws["wb1"], st)
        for i in range(1, self.depth):
            if self.use_tc:
                self.lib.prep_weights_tc(P.conv_w[i], self.ws["wt_f"][i], self.ws["wt_d"][i], st)
            else:
                self.lib.prep_weights(P.conv_w[i], self.ws["wt_f"][i], self.ws["wt_d"][i], st)

    def conv64(self, a, wt, bias, out, partials, B, H, W, st):
        """64->64 3x3 convolution over the padded layout (forward: wt_f + bias + partials; dgrad: wt_d)."""
        if self.use_tc:
            self.lib.conv3x3_tc_fwd(a, wt, bias, out, partials, self.ws["tc_err"], B, H, W, st)
        else:
            self.lib.conv3x3_fwd(a, wt, bias, out, partials, B, H, W, st)

    def check_tc(self):
        if self.use_tc and self.ws is not None and int(self.ws["tc_err"].item()) != 0:
            self.ws["tc_err"].zero_()
            raise RuntimeError("tcgen05 convolution pipeline reported a barrier time-out: activations / gradients of the "
                               "steps since the last check are not trustworthy")

    def forward(self, x, P, ipe, training, update_running=True):
        """x [B,3,H,W] (device, contiguous).  Returns features [B, D] (NHWC-flattened view of a workspace)."""
        B = x.shape[0]
        assert x.is_contiguous() and x.dtype == torch.float32 and B % ipe == 0
        E = B // ipe
        self.ensure(B, E)
        ws, lib, st = self.ws, self.lib, _stream(self.dev)
        self.prepare_weights(P)
        for i, L in enumerate(self.layers):
            H, W, pool = L["H"], L["W"], int(L["pool"])
            last = i == self.depth - 1
            partials = ws["partials"][i] if training else None
            if i == 0 and self.conv1_tc:
                if training:
                    lib.conv1_tc(x, ws["wb1"], P.conv_b[0], ws["y"][0], partials, None, None, None, None, None,
                                 ws["tc_err"], B, H, W, ipe, 0, st)
                else:       # fused conv1 + BatchNorm(running stats) + ReLU + MaxPool: y[0] is never materialised
                    lib.bn_eval_prepare(P.bn_rm[0], P.bn_rv[0], ws["eval_mean"][0], ws["eval_invstd"][0], 64, BN_EPS, st)
                    lib.conv1_tc(x, ws["wb1"], P.conv_b[0], None, None, ws["eval_mean"][0], ws["eval_invstd"][0],
                                 P.bn_w[0], P.bn_b[0], ws["act"][0], ws["tc_err"], B, H, W, 0, 2, st)
                    continue
            elif i == 0:
                lib.conv1_fwd(x, P.conv_w[0], P.conv_b[0], ws["y"][0], partials, B, H, W, st)
            else:
                self.conv64(ws["act"][i - 1], ws["wt_f"][i], P.conv_b[i], ws["y"][i], partials, B, H, W, st)
            if training:
                lib.bn_finalize(partials, B, ws["T"][i], ipe, H * W, ws["mean"][i], ws["invstd"][i],
                                P.bn_rm[i] if update_running else None, P.bn_rv[i] if update_running else None,
                                ws["scratch_d"], BN_MOMENTUM, BN_EPS, st)
                mean, invstd, ipe_arg = ws["mean"][i], ws["invstd"][i], ipe
            else:
                lib.bn_eval_prepare(P.bn_rm[i], P.bn_rv[i], ws["eval_mean"][i], ws["eval_invstd"][i], 64, BN_EPS, st)
                mean, invstd, ipe_arg = ws["eval_mean"][i], ws["eval_invstd"][i], 0
            lib.bn_relu_pool_fwd(ws["y"][i], mean, invstd, P.bn_w[i], P.bn_b[i], ws["act"][i], B, H, W, ipe_arg,
                                 0 if i == 0 else 1, 0 if last else 1, pool, st)
        return ws["act"][-1].view(B, self.D)

    # ------------------------------------------------------------------ backward (after a training forward)
    def backward(self, x, gfeat, P, G, ipe):
        """gfeat [B, D] gradient w.r.t. the features; fills G (a ConvNetParams of gradient tensors)."""
        B = x.shape[0]
        ws, lib, st = self.ws, self.lib, _stream(self.dev)
        gout = gfeat
        for i in range(self.depth - 1, -1, -1):
            L = self.layers[i]
            H, W, pool = L["H"], L["W"], int(L["pool"])
            last = i == self.depth - 1
            lib.bn_relu_pool_bwd(ws["y"][i], gout, ws["mean"][i], ws["invstd"][i], P.bn_w[i], P.bn_b[i], ws["gy"][i],
                                 G.bn_w[i], G.bn_b[i], ws["bwd_partial"], ws["bwd_sums"], ws["scratch_d"], B, H, W, ipe,
                                 0 if i == 0 else 1, 0 if last else 1, pool, st)
            if i == 0 and ws["gy"][0] is None:
                getattr(lib, self.l0_fn)(x, ws["y"][0], gout, ws["mean"][0], ws["invstd"][0], P.bn_w[0], P.bn_b[0],
                                         ws["bwd_sums"], G.conv_w[0], G.conv_b[0], ws["wgrad_scratch"], B, H, W, ipe,
                                         0 if last else 1, st)
            elif i == 0:
                lib.conv1_wgrad(x, ws["gy"][0], G.conv_w[0], G.conv_b[0], ws["wgrad_scratch"], B, H, W, st)
            else:
                if self.use_tc and self.wgrad_tc:
                    lib.conv3x3_wgrad_tc(ws["act"][i - 1], ws["gy"][i], G.conv_w[i], G.conv_b[i], ws["wgrad_scratch"],
                                         ws["tc_err"], B, H, W, st)
                else:
                    lib.conv3x3_wgrad(ws["act"][i - 1], ws["gy"][i], G.conv_w[i], G.conv_b[i], ws["wgrad_scratch"], B,
                                      H, W, st)
                self.conv64(ws["gy"][i], ws["wt_d"][i], None, ws["gact"][i - 1], None, B, H, W, st)
                gout = ws["gact"][i - 1]


class Conv3Engine:
    """Conv3 of the QMUL regression path (reference backbone.py:379-402): three 3x3 / stride 2 / dilation 2 un-padded
    convolutions, each followed by ReLU, then flatten.  NHWC generic-convolution kernels (csrc/conv_generic.cu)."""

    CH = 36

    def __init__(self, lib, device="cuda"):
        self.lib, self.dev = lib, torch.device(device)
        self.key = None

    def _alloc(self, n, H, W):
        lib, dev, f32 = self.lib, self.dev, torch.float32
        sizes = [(H, W)]
        for _ in range(3):
            h, w = sizes[-1]
            sizes.append((lib.conv2d_out_size(h, 3, 2, 0, 2), lib.conv2d_out_size(w, 3, 2, 0, 2)))
        self.sizes = sizes
        self.x_nhwc = torch.empty(n, H, W, 3, device=dev, dtype=f32)
        self.act = [torch.empty(n, h, w, self.CH, device=dev, dtype=f32) for (h, w) in sizes[1:]]
        self.gact = [torch.empty(n, h, w, self.CH, device=dev, dtype=f32) for (h, w) in sizes[1:3]]
        big = 0
        for i in range(3):
            cin = 3 if i == 0 else self.CH
            ho, wo = sizes[i + 1]
            big = max(big, lib.conv2d_wgrad_nsplit(n * ho * wo) * 9 * cin * self.CH)
        self.scratch = torch.empty(big, device=dev, dtype=f32)
        self.D = sizes[3][0] * sizes[3][1] * self.CH
        self.P = sizes[3][0] * sizes[3][1]
        self.key = (n, H, W)

    def forward(self, x, weights, biases):
        """x [n,3,H,W] NCHW -> features [n, D] (NHWC-flattened view)."""
        n, _, H, W = x.shape
        if self.key != (n, H, W):
            self._alloc(n, H, W)
        lib, st = self.lib, _stream(self.dev)
        lib.nchw_to_nhwc(x, self.x_nhwc, n, 3, H, W, st)
        src = self.x_nhwc
        for i in range(3):
            h, w = self.sizes[i]
            lib.conv2d_fwd(src, weights[i], biases[i], self.act[i], n, h, w, 3 if i == 0 else self.CH, self.CH, 3, 3, 2, 0,
                           2, 1, st)
            src = self.act[i]
        return self.act[2].view(n, self.D)

    def backward(self, gfeat, weights, gw, gb):
        """gfeat [n, D]; fills gw[i] / gb[i] (reference layout)."""
        n = self.key[0]
        lib, st = self.lib, _stream(self.dev)
        g = gfeat.view(self.act[2].shape)
        for i in (2, 1, 0):
            h, w = self.sizes[i]
            cin = 3 if i == 0 else self.CH
            src = self.x_nhwc if i == 0 else self.act[i - 1]
            lib.conv2d_wgrad(src, g, self.act[i], gw[i], gb[i], self.scratch, n, h, w, cin, self.CH, 3, 3, 2, 0, 2, 1, st)
            if i > 0:
                lib.conv2d_dgrad(g, self.act[i], weights[i], self.gact[i - 1], n, h, w, cin, self.CH, 3, 3, 2, 0, 2, 1, st)
                g = self.gact[i - 1]


class GPHeadParams:
    """Tensors of the GP head: bn_out (optional) + per-class raw hyper-parameters."""

    def __init__(self):
        self.bn_w = self.bn_b = self.bn_rm = self.bn_rv = None
        self.raw_outputscale = self.constant = self.raw_noise = None
        self.raw_param = None      # per-class raw variance / lengthscale / offset of the kernel family


class GPHead:
    """bn_out -> L2 normalise -> Gram -> C exact GPs, forward/backward/predict over E packed episodes.

    ``kernel`` in {"bncossim", "cossim", "linear"(variance frozen at its value)}: the base kernel is the
    Gram matrix of the (normalised) features, shared by the C one-vs-rest models (DKT.py:366-370).
    """

    FAMILY = {"linear": 0, "rbf": 1, "matern": 2, "poli1": 3, "poli2": 4}

    def __init__(self, lib, kernel, n_way, D, Cch, P, device="cuda"):
        assert kernel in ("bncossim", "cossim") or kernel in self.FAMILY, kernel
        self.lib, self.kernel, self.C, self.D, self.Cch, self.P = lib, kernel, n_way, D, Cch, P
        self.dev = torch.device(device)
        self.bn = kernel == "bncossim"
        self.normalize = kernel in ("bncossim", "cossim")
        # cossim / bncossim: the Gram matrix of the normalised features IS the base kernel, shared by all classes
        # (variance frozen at 1).  Other kernels: per-class epilogue on the Gram of the (centred) features.
        self.family = self.FAMILY.get(kernel)
        self.centred = kernel in ("rbf", "matern")
        self.key = None
        # Gram / cross-kernel products on tcgen05 (csrc/gram_tc.cu) on a GPU; the FFMA kernel serves the host test build
        # and the shapes the tensor-core kernel does not take (fewer than 128 rows in total, D not a multiple of 4)
        self.use_tc = self.dev.type == "cuda" and lib.has("dktb_gram_tc")
        self.tc_err = torch.zeros(1, device=self.dev, dtype=torch.int32) if self.use_tc else None

    def gram(self, x1, x2, out, E, M, N):
        """out [E,M,N] = x1 [E,M,D] . x2 [E,N,D]^T"""
        st = _stream(self.dev)
        if self.use_tc and self.lib.gram_tc_ok(E, M, N, self.D):
            need = self.lib.gram_tc_scratch_floats(E, M, N, self.D)
            if getattr(self, "_gram_scratch", None) is None or self._gram_scratch.numel() < need:
                self._gram_scratch = torch.empty(need, device=self.dev)
            self.lib.gram_tc(x1, x2, out, self._gram_scratch, self.tc_err, E, M, N, self.D, st)
        else:
            self.lib.gram(x1, x2, out, E, M, N, self.D, st)

    def _alloc(self, E, N):
        dev, f32, C, D = self.dev, torch.float32, self.C, self.D
        w = {}
        w["z"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["zh"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["zh_train"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["inv"] = torch.empty(E * N, device=dev, dtype=f32)
        w["bn_mean"] = torch.empty(E, D, device=dev, dtype=f32)
        w["bn_invstd"] = torch.empty(E, D, device=dev, dtype=f32)
        w["bn_var"] = torch.empty(E, D, device=dev, dtype=f32)
        w["gram"] = torch.empty(E, N, N, device=dev, dtype=f32)
        w["alpha"] = torch.empty(E, C, N, device=dev, dtype=f32)
        w["loss_terms"] = torch.empty(E, C, device=dev, dtype=f32)
        w["info"] = torch.zeros(E, C, device=dev, dtype=torch.int32)
        w["info_sticky"] = torch.zeros(2, device=dev, dtype=torch.int32)     # (worst failure, deepest jitter retry) so far
        w["dk"] = torch.empty(E, C, N, N, device=dev, dtype=f32)
        w["dhyper"] = torch.empty(E, C, 3, device=dev, dtype=f32)
        w["loss"] = torch.empty(E, device=dev, dtype=f32)
        w["hyper"] = torch.empty(C, 3, device=dev, dtype=f32)
        w["dzh"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["dz"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["df"] = torch.empty(E, N, D, device=dev, dtype=f32)
        w["pgrad"] = torch.empty(E * 2 * D, device=dev, dtype=f32)
        if self.family is not None:
            w["d2"] = torch.empty(E, N, N, device=dev, dtype=f32)
            w["kb"] = torch.empty(E, C, N, N, device=dev, dtype=f32)
            w["dg"] = torch.empty(E, N, N, device=dev, dtype=f32)
            w["dparam"] = torch.empty(C, device=dev, dtype=f32)
            w["kscratch"] = torch.empty(E * C * N, device=dev, dtype=f32)
        self.w, self.key = w, (E, N)

    def ensure(self, E, N):
        if self.key != (E, N):
            self._alloc(E, N)

    def embed(self, feats, HP, E, N, training, update_running=True, out=None):
        """features [E*N, D] -> (normalised) embedding [E,N,D] (reference: bn_out in the trunk + F.normalize)."""
        lib, st, w = self.lib, _stream(self.dev), self.w
        cur = feats.view(E, N, self.D)
        if self.bn:
            lib.bn1d_fwd(cur, HP.bn_w, HP.bn_b, HP.bn_rm, HP.bn_rv, w["z"], w["bn_mean"], w["bn_invstd"], w["bn_var"],
                         E, N, self.D, self.Cch, self.P, int(training), int(update_running), BN_MOMENTUM, BN_EPS, st)
            cur = w["z"]
        dst = out if out is not None else w["zh"]
        if self.normalize:
            lib.l2norm_fwd(cur, dst, w["inv"] if training else None, E * N, self.D, 1e-12, st)
        else:
            dst.copy_(cur)      # keep the embedding out of the backbone workspace (device memcpy)
        return dst

    def fit(self, zh, targets, HP, E, N, want_grad, grad_scale=1.0, jitter=GP_JITTER):
        """Gram + C Cholesky systems per episode.  targets [C,N] (shared by all episodes)."""
        lib, st, w, C = self.lib, _stream(self.dev), self.w, self.C
        if self.family is None:
            self.gram(zh, zh, w["gram"], E, N, N)
            kb, stride = w["gram"], 0
        else:
            if self.centred:      # rbf / matern: direct squared distances (no Gram cancellation, exact zero diagonal)
                lib.sqdist(zh, zh, w["d2"], E, N, N, self.D, st)
                lib.kernel_fwd(self.family, None, w["d2"], HP.raw_param, w["kb"], E, C, N, N, st)
            else:
                self.gram(zh, zh, w["gram"], E, N, N)
                lib.kernel_fwd(self.family, w["gram"], None, HP.raw_param, w["kb"], E, C, N, N, st)
            kb, stride = w["kb"], N * N
        # one CTA per system entirely in shared memory up to the BASELINE episode size (N = 105); beyond that the tiled
        # kernel on a global workspace is faster (N = 128: 0.57 vs 0.21 ms, N = 165: 0.91 vs 0.36 ms per 160 systems,
        # profiles/r01_gp_size_sweep.txt)
        if N <= min(lib.gp_max_n(), GP_SMEM_CROSSOVER):
            lib.gp_fit(kb, stride, targets, 0, HP.raw_outputscale, HP.constant, HP.raw_noise, w["alpha"], None,
                       w["loss_terms"], w["info"], w["dk"] if want_grad else None, w["dhyper"] if want_grad else None,
                       grad_scale, jitter, E, C, N, st)
        else:               # tiled block factorisation on a global workspace (csrc/gp_large.cu)
            if N > lib.gp_large_max_n():
                raise NotImplementedError("exact-GP systems with N = %d > %d" % (N, lib.gp_large_max_n()))
            need = lib.gp_large_work_floats(E, C, N)
            if w.get("gp_work") is None or w["gp_work"].numel() < need:
                w["gp_work"] = torch.empty(need, device=self.dev)
            lib.gp_fit_large(kb, stride, targets, 0, HP.raw_outputscale, HP.constant, HP.raw_noise, w["alpha"], None,
                             w["loss_terms"], w["info"], w["dk"] if want_grad else None,
                             w["dhyper"] if want_grad else None, w["gp_work"], grad_scale, jitter, E, C, N, st)
        lib.gp_reduce(w["loss_terms"], w["dhyper"] if want_grad else None, w["loss"], w["hyper"] if want_grad else None,
                      E, C, st)
        lib.gp_info_accumulate(w["info"], w["info_sticky"], E * C, st)
        return w["loss"]

    def check(self, what="Cholesky"):
        """One host read-back of the Cholesky status of every fit since the last call: warns like psd_safe_cholesky when a
        system needed jitter, raises when one stayed not positive definite (the reference raises at that step)."""
        hi, lo = self.w["info_sticky"].tolist()
        self.w["info_sticky"].zero_()
        if self.use_tc and int(self.tc_err.item()) != 0:
            self.tc_err.zero_()
            raise RuntimeError("tcgen05 Gram pipeline reported a barrier time-out")
        report_info(hi, lo, what)

    def hyper_grads(self, HP, GH, E, N):
        """After fit(want_grad=True): only the GP hyper-parameter gradients (outputscale, constant, kernel parameter) --
        the features are constants (test-time adaptation, DKT.correct with N > 0)."""
        lib, st, w = self.lib, _stream(self.dev), self.w
        if self.family is not None:
            lib.kernel_bwd(self.family, None if self.centred else w["gram"], w["d2"] if self.centred else None,
                           HP.raw_param, w["dk"], w["dg"], w["dparam"], w["kscratch"], E, self.C, N, st)
            GH.raw_param.copy_(w["dparam"])
        GH.raw_outputscale.copy_(w["hyper"][:, 0])
        GH.constant.copy_(w["hyper"][:, 1])

    def backward(self, feats, zh, HP, GH, E, N):
        """Gradient w.r.t. the backbone features [E*N, D]; fills GH.{bn_w,bn_b,raw_outputscale,constant,raw_param}."""
        lib, st, w = self.lib, _stream(self.dev), self.w
        if self.family is None:
            lib.gram_bwd(w["dk"], zh, w["dzh"], E, self.C, N, self.D, 1.0, st)
        else:
            lib.kernel_bwd(self.family, None if self.centred else w["gram"], w["d2"] if self.centred else None,
                           HP.raw_param, w["dk"], w["dg"], w["dparam"], w["kscratch"], E, self.C, N, st)
            GH.raw_param.copy_(w["dparam"])
            # dg carries the distance gradient both off the diagonal (-2 A / l^2) and on it (row sums), so
            # (dg + dg^T) Z is exactly sum_j A_ij 2 (z_i - z_j) / l^2 -- no cancellation in the backward pass
            lib.gram_bwd(w["dg"], zh, w["dzh"], E, 1, N, self.D, 1.0, st)
        g = w["dzh"]
        if self.normalize:
            lib.l2norm_bwd(zh, g, w["inv"], w["dz"], E * N, self.D, st)
            g = w["dz"]
        if self.bn:
            lib.bn1d_bwd(feats.view(E, N, self.D), g, HP.bn_w, w["bn_mean"], w["bn_invstd"], w["df"], GH.bn_w, GH.bn_b,
                         w["pgrad"], E, N, self.D, self.Cch, self.P, st)
            g = w["df"]
        GH.raw_outputscale.copy_(w["hyper"][:, 0])
        GH.constant.copy_(w["hyper"][:, 1])
        return g.view(E * N, self.D)

    def predict(self, zh_test, zh_train, HP, E, M, N, mean_out, pred_out, kx_buf):
        """Predictive mean [E,C,M] + class arg-max [E,M]; expects alpha of the current fit in the workspace."""
        lib, st = self.lib, _stream(self.dev)
        if self.family is None:
            self.gram(zh_test, zh_train, kx_buf, E, M, N)
            lib.gp_predict(kx_buf, 0, self.w["alpha"], HP.raw_outputscale, HP.constant, mean_out, pred_out, E, self.C,
                           M, N, st)
            return
        dev, f32, C = self.dev, torch.float32, self.C
        t = self.w.setdefault("pred_tmp", {})
        key = (E, M, N)
        if t.get("key") != key:
            t.update(key=key, kx=torch.empty(E, C, M, N, device=dev, dtype=f32))
        if self.centred:
            lib.sqdist(zh_test, zh_train, kx_buf, E, M, N, self.D, st)
            lib.kernel_fwd(self.family, None, kx_buf, HP.raw_param, t["kx"], E, C, M, N, st)
        else:
            self.gram(zh_test, zh_train, kx_buf, E, M, N)
            lib.kernel_fwd(self.family, kx_buf, None, HP.raw_param, t["kx"], E, C, M, N, st)
        lib.gp_predict(t["kx"], M * N, self.w["alpha"], HP.raw_outputscale, HP.constant, mean_out, pred_out, E, C, M, N,
                       st)


def make_targets(n_way, per_class, device):
    """methods/DKT.py:129-136 / 227-234: one +-1 target vector per class, class-major sample order."""
    t = -torch.ones(n_way, n_way * per_class, dtype=torch.float32)
    for c in range(n_way):
        t[c, c * per_class:(c + 1) * per_class] = 1.0
    return t.to(device)


def report_info(hi, lo, what="Cholesky"):
    """hi > 0: a system stayed not positive definite after the jitter retries -> RuntimeError (GPyTorch 1.0.1 re-raises
    torch's error, later versions raise NotPSDError); lo < 0: -lo retries were needed -> RuntimeWarning with GPyTorch's text."""
    import warnings
    if hi > 0:
        raise RuntimeError("NotPSDError: %s: matrix not positive definite after adding jitter up to %g (first failing "
                           "pivot %d)" % (what, GP_JITTER * 100, hi))
    if lo < 0:
        warnings.warn("A not p.d., added jitter of %.1e to the diagonal" % (GP_JITTER * 10 ** (-lo - 1)), RuntimeWarning)


def check_info(info, what="Cholesky"):
    """Status array of ONE fit (host sync)."""
    if info.numel() == 0:
        return
    report_info(int(info.max().item()), int(info.min().item()), what)


def adam_hparams():
    return dict(beta1=0.9, beta2=0.999, eps=1e-8)


def softplus_inv(x):
    return x + math.log(-math.expm1(-x))
