"""ResNet10/18/34/50/101 feature extractors (reference backbone.py:135-247 SimpleBlock / BottleneckBlock, 330-376
ResNet) over packed episodes: NHWC fp32, generic convolution kernels (csrc/conv_generic.cu) + channel-generic
BatchNorm / pooling kernels (csrc/resnet_ops.cu), every BatchNorm batch = one episode.  The forward records a tape of
(op, tensors); the backward walks it in reverse and accumulates gradients where an activation feeds two consumers
(block input -> main branch and shortcut).  Parameters / gradients are read straight from the modules (their storage
is the flat buffer of methods/DKT.py)."""
import torch

from .engine import BN_EPS, BN_MOMENTUM, _stream


class ResNetEngine:
    def __init__(self, lib, net, device):
        self.lib, self.net, self.dev = lib, net, torch.device(device)
        self.tape = []
        self.keep_tape = False       # tests: keep the consumed tape in `last_tape` (same-branch gradient replay)
        self.last_tape = None
        self.trace = None            # tests may set a list: backward appends (record, gy, gx, gres) per op
        self.D = net.final_feat_dim
        self.P = 1
        self._mma_w = {}             # conv module -> (wf, wd): k-contiguous weight copies for the tensor-core kernels
        self._pads = {}              # (role, B, H, W, C) -> persistent zero-bordered padded-flat buffer
        # tcgen05 / TMA kernel for every stride-1 3x3 and 1x1 convolution whose channel counts are multiples of 64
        # (csrc/conv_tcg.cu); the remaining layers (stem, stride-2 convolutions) run on the mma.sync tiles
        self.use_tcg = self.dev.type == "cuda" and lib.has("dktb_conv_tcg")
        self.tc_err = torch.zeros(1, device=self.dev, dtype=torch.int32) if self.use_tcg else None

    # ------------------------------------------------------------------ helpers
    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def check_tc(self):
        if self.use_tcg and int(self.tc_err.item()) != 0:
            self.tc_err.zero_()
            raise RuntimeError("tcgen05 convolution pipeline reported a barrier time-out: activations / gradients of the "
                               "steps since the last check are not trustworthy")

    def _padbuf(self, role, B, H, W, C):
        """Persistent padded-flat buffer [B, H+2, W+2, C]: the border is zeroed once and never written again (pad_copy and
        the convolution only touch the interior)."""
        key = (role, B, H, W, C)
        if key not in self._pads:
            self._pads[key] = torch.zeros(B, H + 2, W + 2, C, device=self.dev)
        return self._pads[key]

    # Activations / gradients of the tcgen05 3x3 layers live in padded-flat buffers [B, H+2, W+2, C]; the engine hands
    # their INTERIOR VIEW around (logical shape [B, H, W, C], non-contiguous), the kernels get the buffer base plus a layout
    # bit per tensor (dktb_*_l entry points), so no copy separates a convolution from the BatchNorm around it.
    @staticmethod
    def _is_pad(t):
        return t is not None and t.dim() == 4 and not t.is_contiguous()

    @staticmethod
    def _base(t):
        return t._base if ResNetEngine._is_pad(t) else t

    def _newpad(self, B, H, W, C, zero=True):
        """Interior view of a fresh padded-flat buffer; zero: the border is cleared (needed when a convolution reads it)."""
        buf = self._new(B, H + 2, W + 2, C)
        if zero:
            self.lib.zero_border(buf, B, H, W, C, _stream(self.dev))
        return buf[:, 1:-1, 1:-1, :]

    def _like(self, t):
        """Gradient buffer in the layout of its activation."""
        if self._is_pad(t):
            return self._newpad(*t.shape)
        return torch.empty(t.shape, device=self.dev, dtype=t.dtype)

    def _dense(self, t):
        if not self._is_pad(t):
            return t
        B, H, W, C = t.shape
        d = self._new(B, H, W, C)
        self.lib.pad_copy(d, t._base, B, H, W, C, 1, _stream(self.dev))
        return d

    def _padded_base(self, t, role):
        """Padded-flat buffer holding t: its own one, or a persistent scratch buffer filled by pad_copy."""
        if self._is_pad(t):
            return t._base
        B, H, W, C = t.shape
        buf = self._padbuf(role, B, H, W, C)
        self.lib.pad_copy(t, buf, B, H, W, C, 0, _stream(self.dev))
        return buf

    def _lay(self, x=None, y=None, gy=None, gx=None, res=None):
        return (1 if self._is_pad(x) else 0) | (2 if self._is_pad(y) else 0) | (4 if self._is_pad(gy) else 0) | \
            (8 if self._is_pad(gx) else 0) | (16 if self._is_pad(res) else 0)

    def _tcg_weights(self, m):
        key = ("tcg", id(m))
        if key not in self._mma_w:
            n = self.lib.conv_tcg_weight_floats(m.in_channels, m.out_channels, m.kernel_size[0])
            self._mma_w[key] = (self._new(n), self._new(n))
        return self._mma_w[key]

    def _tcg_ok(self, m, W):
        return self.use_tcg and self.lib.conv_tcg_ok(m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0],
                                                     m.padding[0], m.dilation[0], W)

    def _s2_ok(self, m, H, W):
        """Stride-2 3x3 / pad 1 layer that runs on the tcgen05 kernel over the space-to-depth input?"""
        return (self.use_tcg and self.lib.has("dktb_conv_tcg_s2") and m.kernel_size[0] == 3 and m.stride[0] == 2
                and m.padding[0] == 1 and m.dilation[0] == 1
                and self.lib.conv_tcg_s2_ok(m.in_channels, m.out_channels, H, W))

    def _sub2_ok(self, m, H, W):
        """Stride-2 1x1 shortcut that runs as a gather + tcgen05 GEMM?"""
        return (self.use_tcg and self.lib.has("dktb_subsample2") and m.kernel_size[0] == 1 and m.stride[0] == 2
                and m.padding[0] == 0 and H % 2 == 0 and W % 2 == 0
                and self.lib.conv_tcg_ok(m.in_channels, m.out_channels, 1, 1, 0, 1, W // 2))

    def _s2_weights(self, m):
        key = ("tcg_s2", id(m))
        if key not in self._mma_w:
            n = self.lib.conv_tcg_s2_weight_floats(m.in_channels, m.out_channels)
            self._mma_w[key] = (self._new(n), self._new(n))
        return self._mma_w[key]

    def _conv(self, x, m, relu=0):
        B, H, W, Cin = x.shape
        R = m.kernel_size[0]
        st, pad, dil = m.stride[0], m.padding[0], m.dilation[0]
        Ho = self.lib.conv2d_out_size(H, R, st, pad, dil)
        Wo = self.lib.conv2d_out_size(W, R, st, pad, dil)
        bias = m.bias.data if m.bias is not None else None
        if not relu and self._tcg_ok(m, W):
            # tcgen05: 3x3 over padded-flat buffers (the zero border is the padding), 1x1 as a GEMM over the dense rows
            Cout, sm = m.out_channels, _stream(self.dev)
            wf, wd = self._tcg_weights(m)
            self.lib.prep_weights_tcg(m.weight.data, wf, wd, Cout, Cin, R, sm)
            if R == 3:
                out = self._newpad(B, H, W, Cout, zero=False)        # only its interior is ever read (BatchNorm follows)
                self.lib.conv_tcg(self._padded_base(x, "in"), wf, bias, out._base, self.tc_err, B, H, W, Cin, Cout, R, sm)
            else:
                out = self._new(B, Ho, Wo, Cout)
                self.lib.conv_tcg(self._dense(x), wf, bias, out, self.tc_err, B, H, W, Cin, Cout, R, sm)
            self.tape.append(("conv", x, out, m, (B, H, W, Cin, R, st, pad, dil)))
            return out
        if not relu and self._s2_ok(m, H, W):
            # stride-2 3x3: the same tcgen05 kernel over the space-to-depth input (4 parity planes as 4 Cin channels)
            Cout, sm = m.out_channels, _stream(self.dev)
            wf, wd = self._s2_weights(m)
            self.lib.prep_weights_tcg_s2(m.weight.data, wf, wd, Cout, Cin, sm)
            xs = self._padbuf("s2", B, Ho, Wo, 4 * Cin)
            self.lib.s2d(self._base(x), xs, B, H, W, Cin, int(self._is_pad(x)), 0, sm)
            out = self._newpad(B, Ho, Wo, Cout, zero=False)
            self.lib.conv_tcg_s2(xs, wf, bias, out._base, self.tc_err, B, Ho, Wo, Cin, Cout, 0, sm)
            self.tape.append(("conv", x, out, m, (B, H, W, Cin, R, st, pad, dil)))
            return out
        if not relu and self._sub2_ok(m, H, W):
            # stride-2 1x1 shortcut: every second pixel gathered into a dense tensor, then the tcgen05 GEMM
            Cout, sm = m.out_channels, _stream(self.dev)
            wf, wd = self._tcg_weights(m)
            self.lib.prep_weights_tcg(m.weight.data, wf, wd, Cout, Cin, 1, sm)
            xg = self._new(B, Ho, Wo, Cin)
            self.lib.subsample2(self._base(x), xg, B, H, W, Cin, int(self._is_pad(x)), 0, sm)
            out = self._new(B, Ho, Wo, Cout)
            self.lib.conv_tcg(xg, wf, bias, out, self.tc_err, B, Ho, Wo, Cin, Cout, 1, sm)
            self.tape.append(("conv", x, out, m, (B, H, W, Cin, R, st, pad, dil)))
            return out
        out = self._new(B, Ho, Wo, m.out_channels)
        xd = self._dense(x)
        if not relu and self.lib.conv2d_mma_ok(Cin, m.out_channels):
            # tensor-core tiles (mma.sync 3xTF32): refresh the two k-contiguous weight copies, then forward from `wf`;
            # the backward of the same step reads `wd`
            wf, wd = self._mma_weights(m)
            self.lib.conv2d_prep_mma(m.weight.data, wf, wd, m.out_channels, Cin, R, R, _stream(self.dev))
            self.lib.conv2d_fwd_mma(xd, wf, bias, out, B, H, W, Cin, m.out_channels, R, R, st, pad, dil, _stream(self.dev))
        elif not relu and Cin < 16 and self.lib.conv2d_mma_ok(32, 32):
            # the stem: reduction over the flattened (r, s, ci) index
            key = ("flat", id(m))
            if key not in self._mma_w:
                self._mma_w[key] = self._new(m.out_channels * self.lib.conv2d_flat_k(Cin, R, R))
            self.lib.conv2d_prep_flat_mma(m.weight.data, self._mma_w[key], m.out_channels, Cin, R, R, _stream(self.dev))
            self.lib.conv2d_fwd_flat_mma(xd, self._mma_w[key], bias, out, B, H, W, Cin, m.out_channels, R, R, st, pad, dil,
                                         _stream(self.dev))
        else:
            self.lib.conv2d_fwd(xd, m.weight.data, bias, out, B, H, W, Cin, m.out_channels, R, R, st, pad, dil, relu,
                                _stream(self.dev))
        self.tape.append(("conv", x, out, m, (B, H, W, Cin, R, st, pad, dil)))
        return out

    def _stem(self, x, xh, m):
        """7x7 / stride 2 stem: the tcgen05 kernel reads the NCHW batch directly (csrc/stem_tc.cu); the tape keeps the NHWC
        copy for the weight gradient.  Other stems (different geometry) take the generic path."""
        B, _, H, W = x.shape
        R, sv, pad, dil = m.kernel_size[0], m.stride[0], m.padding[0], m.dilation[0]
        if not (self.use_tcg and self.lib.has("dktb_stem_tc")
                and self.lib.stem_tc_ok(m.in_channels, m.out_channels, R, sv, pad, dil, H, W)):
            return self._conv(xh, m)
        key = ("stem", id(m))
        if key not in self._mma_w:
            self._mma_w[key] = self._new(self.lib.stem_tc_weight_floats())
        sm = _stream(self.dev)
        self.lib.prep_weights_stem_tc(m.weight.data, self._mma_w[key], sm)
        out = self._new(B, H // 2, W // 2, m.out_channels)
        self.lib.stem_tc(x, self._mma_w[key], m.bias.data if m.bias is not None else None, out, self.tc_err, B, H, W, sm)
        self.tape.append(("conv", xh, out, m, (B, H, W, m.in_channels, R, sv, pad, dil)))
        self._stem_nchw = (xh.data_ptr(), x)       # the weight-gradient kernel reads the NCHW batch as well
        return out

    def _mma_weights(self, m):
        key = id(m)
        if key not in self._mma_w:
            n = m.weight.numel()
            self._mma_w[key] = (self._new(n), self._new(n))
        return self._mma_w[key]

    def _bn(self, x, m, ipe, training, res=None, relu=0, pad_out=False):
        """pad_out: write the output into a zero-bordered padded-flat buffer (its consumer is a tcgen05 3x3 layer)."""
        B, H, W, C = x.shape
        st = _stream(self.dev)
        E = B // ipe if training else 1
        mean, invstd = self._new(E, C), self._new(E, C)
        y = self._newpad(B, H, W, C) if pad_out else self._new(B, H, W, C)
        lay = self._lay(x=x, y=y, res=res)
        if training:
            partial = self._new(self.lib.bn2d_partial_floats(B, H * W, C))
            self.lib.bn2d_stats_l(self._base(x), mean, invstd, m.running_mean, m.running_var, partial, B, H * W, C, ipe,
                                  BN_MOMENTUM, BN_EPS, W, lay, st)
        else:
            self.lib.bn_eval_prepare(m.running_mean, m.running_var, mean, invstd, C, BN_EPS, st)
        self.lib.bn2d_apply_l(self._base(x), mean, invstd, m.weight.data, m.bias.data, self._base(res), self._base(y), B,
                              H * W, C, ipe if training else 0, relu, W, lay, st)
        self.tape.append(("bn", x, y, m, mean, invstd, res, relu, ipe))
        return y

    def _feeds_tcg3(self, conv, W):
        """Is `conv` a 3x3 layer of the tcgen05 kernel at row width W (then its input should be produced padded)?"""
        return conv.kernel_size[0] == 3 and self._tcg_ok(conv, W)

    # ------------------------------------------------------------------ forward
    def forward(self, x, ipe, training):
        """x [B,3,H,W] NCHW -> features [B, D]."""
        lib, st, net = self.lib, _stream(self.dev), self.net
        self.tape = []
        self._stem_nchw = None
        B, _, H, W = x.shape
        xh = self._new(B, H, W, 3)
        lib.nchw_to_nhwc(x, xh, B, 3, H, W, st)
        t = net.trunk
        out = self._stem(x, xh, t[0])
        out = self._bn(out, t[1], ipe, training, relu=1)
        Bq, Hq, Wq, Cq = out.shape
        Ho, Wo = (Hq + 2 - 3) // 2 + 1, (Wq + 2 - 3) // 2 + 1
        pooled = self._new(Bq, Ho, Wo, Cq)
        idx = self._new(Bq, Ho, Wo, Cq, dtype=torch.uint8)
        lib.maxpool3_fwd(out, pooled, idx, Bq, Hq, Wq, Cq, st)
        self.tape.append(("maxpool", out, pooled, idx))
        out = pooled
        blocks = list(net.blocks())
        for bi, blk in enumerate(blocks):
            nxt = blocks[bi + 1] if bi + 1 < len(blocks) else None
            if blk.kind == "simple":
                o = self._conv(out, blk.C1)
                o = self._bn(o, blk.BN1, ipe, training, relu=1, pad_out=self._feeds_tcg3(blk.C2, o.shape[2]))
                o = self._conv(o, blk.C2)
                if blk.shortcut_type == "identity":
                    sh = out
                else:
                    sh = self._bn(self._conv(out, blk.shortcut), blk.BNshortcut, ipe, training, relu=0)
                out = self._bn(o, blk.BN2, ipe, training, res=sh, relu=1,
                               pad_out=nxt is not None and self._feeds_tcg3(nxt.C1, o.shape[2]))
            else:
                sh = out if blk.shortcut_type == "identity" else self._conv(out, blk.shortcut)
                o = self._conv(out, blk.C1)
                o = self._bn(o, blk.BN1, ipe, training, relu=1, pad_out=self._feeds_tcg3(blk.C2, o.shape[2]))
                o = self._bn(self._conv(o, blk.C2), blk.BN2, ipe, training, relu=1)
                out = self._bn(self._conv(o, blk.C3), blk.BN3, ipe, training, res=sh, relu=1)
        Bq, Hq, Wq, Cq = out.shape
        feats = self._new(Bq, Cq)
        lib.avgpool_fwd(out, feats, Bq, Hq * Wq, Cq, st)
        self.tape.append(("avgpool", out, feats))
        return feats

    # ------------------------------------------------------------------ backward
    def backward(self, gfeat):
        """gfeat [B, D]; fills ``.grad`` of every backbone parameter (views into the flat gradient buffer)."""
        lib, st = self.lib, _stream(self.dev)
        grads = {}

        def give(t, g):
            k = t.data_ptr()
            if k not in grads:
                grads[k] = g
            elif g.dim() == 4:
                a = grads[k]
                lib.add_inplace_l(self._base(a), self._base(g), g.shape[0], g.shape[1] * g.shape[2], g.shape[3], g.shape[2],
                                  self._lay(x=a, y=g), st)
            else:
                lib.add_inplace(grads[k], g, g.numel(), st)

        last = self.tape[-1]
        give(last[2], gfeat.contiguous())
        for rec in reversed(self.tape):
            kind = rec[0]
            if kind == "avgpool":
                _, x, y = rec
                gx = self._new(*x.shape)
                lib.avgpool_bwd(grads.pop(y.data_ptr()), gx, x.shape[0], x.shape[1] * x.shape[2], x.shape[3], st)
                give(x, gx)
            elif kind == "maxpool":
                _, x, y, idx = rec
                gx = self._new(*x.shape)
                lib.maxpool3_bwd(self._dense(grads.pop(y.data_ptr())), idx, gx, x.shape[0], x.shape[1], x.shape[2], x.shape[3], st)
                give(x, gx)
            elif kind == "bn":
                _, x, y, m, mean, invstd, res, relu, ipe = rec
                B, H, W, C = x.shape
                gy = grads.pop(y.data_ptr())
                gx = self._like(x)
                gres = self._like(res) if res is not None else None
                partial = self._new(self.lib.bn2d_partial_floats(B, H * W, C))
                sums = self._new((B // ipe) * C * 2)
                lib.bn2d_bwd_l(self._base(x), self._base(y), self._base(gy), mean, invstd, m.weight.data, self._base(gx),
                               self._base(gres), m.weight.grad, m.bias.grad, partial, sums, B, H * W, C, ipe, relu, W,
                               self._lay(x=x, y=y, gy=gy, gx=gx, res=gres), st)
                if self.trace is not None:
                    self.trace.append((rec, gy, gx.clone(), None if gres is None else gres.clone()))
                give(x, gx)
                if res is not None:
                    give(res, gres)
            elif kind == "conv":
                _, x, y, m, (B, H, W, Cin, R, stv, pad, dil) = rec
                gy = grads.pop(y.data_ptr())
                Cout = m.out_channels
                gx = None
                if self._tcg_ok(m, W):
                    # tcgen05 weight gradient and dgrad (the forward kernel on the flipped / transposed weights); 3x3:
                    # both operands in the persistent zero-bordered padded-flat buffers, 1x1: straight on the dense rows
                    gx = self._like(x)
                    wd = self._tcg_weights(m)[1]
                    scratch = self._new(lib.wgrad_tcg_scratch_floats(B, H, W, Cin, Cout, R))
                    bg = m.bias.grad if m.bias is not None else None
                    if R == 3:
                        xp, gyp = self._padded_base(x, "in"), self._padded_base(gy, "out")
                        lib.wgrad_tcg(xp, gyp, m.weight.grad, bg, scratch, self.tc_err, B, H, W, Cin, Cout, R, st)
                        if self._is_pad(gx):
                            lib.conv_tcg(gyp, wd, None, gx._base, self.tc_err, B, H, W, Cout, Cin, R, st)
                        else:           # dense input: through the scratch buffer (free again after the weight gradient)
                            xp = self._padbuf("in", B, H, W, Cin)
                            lib.conv_tcg(gyp, wd, None, xp, self.tc_err, B, H, W, Cout, Cin, R, st)
                            lib.pad_copy(gx, xp, B, H, W, Cin, 1, st)
                    else:
                        xd, gyd = self._dense(x), self._dense(gy)
                        gxd = gx if not self._is_pad(gx) else self._new(B, H, W, Cin)
                        lib.wgrad_tcg(xd, gyd, m.weight.grad, bg, scratch, self.tc_err, B, H, W, Cin, Cout, R, st)
                        lib.conv_tcg(gyd, wd, None, gxd, self.tc_err, B, H, W, Cout, Cin, R, st)
                        if gxd is not gx:
                            lib.pad_copy(gxd, gx._base, B, H, W, Cin, 0, st)
                    if self.trace is not None:
                        self.trace.append((rec, gy, gx.clone(), m.weight.grad.clone()))
                    give(x, gx)
                    continue
                ns = lib.conv2d_wgrad_nsplit(y.shape[0] * y.shape[1] * y.shape[2])
                scratch = self._new(ns * R * R * Cin * Cout)
                xn = getattr(self, "_stem_nchw", None)
                if Cin == 3 and m.bias is None and xn is not None and xn[0] == x.data_ptr() and \
                        lib.has("dktb_stem_wgrad_tc") and not self._is_pad(gy):
                    # the stem's weight gradient on tcgen05 (no input gradient: x is the data)
                    scr = self._new(lib.stem_wgrad_tc_scratch_floats(B, H))
                    lib.stem_wgrad_tc(xn[1], gy, m.weight.grad, scr, self.tc_err, B, H, W, st)
                    if self.trace is not None:
                        self.trace.append((rec, gy, None, m.weight.grad.clone()))
                    continue
                if self._sub2_ok(m, H, W):
                    # stride-2 1x1 shortcut: GEMMs on the gathered input, gradient scattered back (zeros at skipped pixels)
                    Ho, Wo = y.shape[1], y.shape[2]
                    xg, gyd = self._new(B, Ho, Wo, Cin), self._dense(gy)
                    lib.subsample2(self._base(x), xg, B, H, W, Cin, int(self._is_pad(x)), 0, st)
                    scr = self._new(lib.wgrad_tcg_scratch_floats(B, Ho, Wo, Cin, Cout, 1))
                    lib.wgrad_tcg(xg, gyd, m.weight.grad, m.bias.grad if m.bias is not None else None, scr, self.tc_err, B,
                                  Ho, Wo, Cin, Cout, 1, st)
                    lib.conv_tcg(gyd, self._tcg_weights(m)[1], None, xg, self.tc_err, B, Ho, Wo, Cout, Cin, 1, st)
                    gx = self._like(x)
                    lib.subsample2(self._base(gx), xg, B, H, W, Cin, int(self._is_pad(gx)), 1, st)
                    if self.trace is not None:
                        self.trace.append((rec, gy, gx.clone(), m.weight.grad.clone()))
                    give(x, gx)
                    continue
                if self._s2_ok(m, H, W):
                    # tcgen05: input gradient dy -> gradient of the space-to-depth input -> unpacked into gx's layout; weight
                    # gradient from the (re-packed) space-to-depth input where its four-plane staging fits shared memory
                    Ho, Wo = y.shape[1], y.shape[2]
                    xs, gyp = self._padbuf("s2", B, Ho, Wo, 4 * Cin), self._padded_base(gy, "out")
                    bg = m.bias.grad if m.bias is not None else None
                    if lib.wgrad_tcg_s2_ok(Cin, Cout, H, W):
                        lib.s2d(self._base(x), xs, B, H, W, Cin, int(self._is_pad(x)), 0, st)
                        scr = self._new(lib.wgrad_tcg_s2_scratch_floats(B, Ho, Wo, Cin, Cout))
                        lib.wgrad_tcg_s2(xs, gyp, m.weight.grad, bg, scr, self.tc_err, B, Ho, Wo, Cin, Cout, st)
                    else:
                        lib.conv2d_wgrad(self._dense(x), self._dense(gy), None, m.weight.grad, bg, scratch, B, H, W, Cin, Cout,
                                         R, R, stv, pad, dil, 0, st)
                    lib.conv_tcg_s2(gyp, self._s2_weights(m)[1], None, xs, self.tc_err, B, Ho, Wo, Cin, Cout, 1, st)
                    gx = self._like(x)
                    lib.s2d(self._base(gx), xs, B, H, W, Cin, int(self._is_pad(gx)), 1, st)
                    if self.trace is not None:
                        self.trace.append((rec, gy, gx.clone(), m.weight.grad.clone()))
                    give(x, gx)
                    continue
                xd, gy = self._dense(x), self._dense(gy)
                lib.conv2d_wgrad(xd, gy, None, m.weight.grad, m.bias.grad if m.bias is not None else None, scratch, B, H, W,
                                 Cin, Cout, R, R, stv, pad, dil, 0, st)
                if Cin > 3:      # no input gradient for the stem
                    gx = self._new(B, H, W, Cin)
                    if lib.conv2d_mma_ok(Cin, Cout):      # `wd` was refreshed by this step's forward (weights unchanged since)
                        lib.conv2d_dgrad_mma(gy, self._mma_weights(m)[1], gx, B, H, W, Cin, Cout, R, R, stv, pad, dil, st)
                    else:
                        lib.conv2d_dgrad(gy, None, m.weight.data, gx, B, H, W, Cin, Cout, R, R, stv, pad, dil, 0, st)
                if self.trace is not None:
                    self.trace.append((rec, gy, None if gx is None else gx.clone(), m.weight.grad.clone()))
                if gx is not None:
                    give(x, gx)
        if self.keep_tape:
            self.last_tape = self.tape
        self.tape = []
