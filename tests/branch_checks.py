"""Same-branch gradient parity for the ConvNet (Conv4 / Conv6) meta-train step at the BASELINE shapes.

Why: a ReLU / max-pool network is piecewise linear.  At the BASELINE shapes (~1e8 gates per step) a handful of
pre-activations sit within fp32 rounding of their threshold (or of a tie inside a 2x2 pool window); which side they
fall on differs between ANY two fp32 evaluations (cuDNN vs MKL, fp32 vs fp64, ...).  One flipped gate changes one term
of a weight-gradient sum of n terms by its full size, i.e. by ~1/sqrt(n) of the result (1e-3 .. 5e-3 at n = 1e5 .. 1e6),
so raw gradients of two correct implementations agree to ~1e-3 only.  What CAN be held to 1e-4 is the gradient on a
FIXED branch.  The device's gates are reproducible bit for bit on the host from the tensors it keeps (pre-BN conv
output y, per-episode mean / invstd, gamma, beta): csrc/dktb_common.cuh ``bn_bwd_route`` evaluates
z = max(fmaf(y - m, gamma * invstd, beta), 0) and takes the first strict maximum of the window; a float64 product-sum
rounded once to float32 is that fmaf.  The check then

  1. replays the WHOLE step (conv trunk -> per-episode BatchNorm -> gates -> bn_out -> normalise -> exact-GP -mll) in
     float64 with the device's gates forced, asserting that every gate that differs from the free-running float64
     choice sits on a pre-activation within 1e-4 of the threshold / tie (i.e. the device branch is a legitimate one);
  2. requires every parameter gradient of the device to match that replay to 1e-4 (north_star), no envelope.

Test infrastructure only."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import backbone as obb
from oracle import episode as oep
from oracle import gp as ogp


def fma32(a, b, c):
    """fmaf(a, b, c) for float32 tensors: exact product and sum in float64, one rounding to float32."""
    return (a.double() * b.double() + c.double()).float()


def device_gates(y, mean, invstd, gamma, beta, ipe, pool):
    """y [B,H,W,64] float32 (NHWC interior of the device's pre-BN conv output), mean / invstd [E,64], gamma / beta [64]
    -> (arg [B,Ho,Wo,64] int64 in scan order dy*2+dx, passed [B,Ho,Wo,64] bool): the gates ``bn_bwd_route`` takes."""
    B, H, W, C = y.shape
    e = torch.arange(B) // ipe
    m = mean[e].view(B, 1, 1, C)
    sc = (gamma.view(1, C) * invstd)[e].view(B, 1, 1, C)          # float32 product, as the kernels form it
    z = fma32(y - m, sc.expand_as(y), beta.view(1, 1, 1, C).expand_as(y)).clamp_min_(0.0)
    if not pool:
        return torch.zeros(B, H, W, C, dtype=torch.long), z > 0
    Ho, Wo = H // 2, W // 2
    win = z[:, :2 * Ho, :2 * Wo].reshape(B, Ho, 2, Wo, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(B, Ho, Wo, C, 4)
    best, arg = win.max(-1)                                        # first maximal value on ties == strict '>' scan
    return arg, best > 0


def collect_device_gates(model, snap, ipe):
    """After a training forward/backward: the gates of every ConvBlock from the engine's workspace."""
    eng = model.feature._engine
    ws = eng.ws
    gates = []
    for i, L in enumerate(eng.layers):
        H, W = L["H"], L["W"]
        y = ws["y"][i].detach().cpu()
        if i > 0:
            y = y[:, 1:H + 1, 1:W + 1, :]
        gates.append(device_gates(y.contiguous(), ws["mean"][i].cpu(), ws["invstd"][i].cpu(),
                                  snap["trunk.%d.BN.weight" % i].float(), snap["trunk.%d.BN.bias" % i].float(), ipe,
                                  L["pool"]))
    return gates


def replay_step64(xs, snap, gp, gates, kernel, n_way, n_support, stats, tie_tol=1e-4):
    """Float64 replay of one packed meta-train step with forced gates.  xs [E,C,SQ,3,H,W]; snap / gp: pre-step
    parameters (any float dtype).  Returns (loss [E], grads keyed like the oracle's)."""
    E, C, SQ = xs.shape[0], xs.shape[1], xs.shape[2]
    N = C * SQ
    P = {k: (v.detach().double().clone().requires_grad_(v.is_floating_point() and (k.endswith(".weight") or k.endswith(".bias")))
             if v.is_floating_point() else v.clone()) for k, v in snap.items()}
    G = {k: v.detach().double().clone() for k, v in gp.items()}
    names = ogp.trainable_gp_names(kernel)
    for k in names:
        G[k].requires_grad_(True)
    depth = len(gates)
    x = xs.reshape(E * N, *xs.shape[3:]).double()
    out = x
    for i in range(depth):
        arg, passed = gates[i]
        y = F.conv2d(out, P["trunk.%d.C.weight" % i], P["trunk.%d.C.bias" % i], padding=1)
        z = torch.cat([F.batch_norm(y[e * N:(e + 1) * N], None, None, P["trunk.%d.BN.weight" % i],
                                    P["trunk.%d.BN.bias" % i], True, 0.0, obb.BN_EPS) for e in range(E)])
        B, Cc, H, W = z.shape
        pool = i < 4
        if pool:
            Ho, Wo = H // 2, W // 2
            win = z[:, :, :2 * Ho, :2 * Wo].reshape(B, Cc, Ho, 2, Wo, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, Cc, Ho, Wo, 4)
            a = arg.permute(0, 3, 1, 2).unsqueeze(-1)                                  # [B,C,Ho,Wo,1]
            sel = win.gather(-1, a).squeeze(-1)
            free = win.detach().clamp_min(0).max(-1).values
        else:
            sel, free = z, z.detach().clamp_min(0)
        pm = passed.permute(0, 3, 1, 2)
        forced = sel * pm
        # legitimacy of the device branch: wherever its output differs from the free-running one, the pre-activation is
        # within tie_tol (relative to the layer's scale) of the threshold / of the window's maximum
        diff = (forced.detach() - free).abs()
        scale = float(free.max())
        nflip = int((diff > 0).sum())
        stats["gates"] = stats.get("gates", 0) + forced.numel()
        stats["flips"] = stats.get("flips", 0) + nflip
        if nflip:
            worst = float(diff.max()) / scale
            stats["worst_flip"] = max(stats.get("worst_flip", 0.0), worst)
            assert worst <= tie_tol, "layer %d: device gate differs off a tie / threshold (%.3e of scale)" % (i, worst)
        out = forced
    feats = out.reshape(out.size(0), -1)
    targets = oep.make_targets(C, SQ, torch.float64)
    losses = []
    pp = dict(P)
    for e in range(E):
        f = feats[e * N:(e + 1) * N]
        if "trunk.bn_out.weight" in pp:
            f = F.batch_norm(f, None, None, pp["trunk.bn_out.weight"], pp["trunk.bn_out.bias"], True, 0.0, obb.BN_EPS)
        if kernel in oep.NORMALIZED_KERNELS:
            f = F.normalize(f, p=2, dim=1)
        losses.append(ogp.mll_loss(kernel, f, targets, G))
    (sum(losses) / E).backward()
    grads = {k: v.grad for k, v in P.items() if torch.is_tensor(v) and v.requires_grad}
    grads.update({k: G[k].grad for k in names})
    return torch.stack([l.detach() for l in losses]), grads, feats.detach()


def rel_err(a, b):
    a, b = a.detach().cpu().double().reshape(-1), b.detach().cpu().double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def device_grads(model, kernel):
    """Device gradients keyed like the oracle's."""
    out = {}
    for i, b in enumerate(model.feature.blocks()):
        out["trunk.%d.C.weight" % i] = b.C.weight.grad
        out["trunk.%d.C.bias" % i] = b.C.bias.grad
        out["trunk.%d.BN.weight" % i] = b.BN.weight.grad
        out["trunk.%d.BN.bias" % i] = b.BN.bias.grad
    bn = getattr(model.feature.trunk, "bn_out", None)
    if bn is not None:
        out["trunk.bn_out.weight"], out["trunk.bn_out.bias"] = bn.weight.grad, bn.bias.grad
    ms = model.model.models
    out["raw_outputscale"] = torch.stack([m.covar_module.raw_outputscale.grad.view(()) for m in ms])
    out["constant"] = torch.stack([m.mean_module.constant.grad.view(()) for m in ms])
    for nm in ("raw_variance", "raw_lengthscale", "raw_offset"):
        if nm in ogp.trainable_gp_names(kernel):
            out[nm] = torch.stack([getattr(m.covar_module.base_kernel, nm).grad.view(()) for m in ms])
    return {k: v.detach().cpu() for k, v in out.items()}
