"""Per-kernel parity checks shared by the CPU-emulation tests (tests/test_kernels_emu.py, host logic of the
same kernel sources) and the GPU tests proper (tests/test_kernels_gpu.py, through the C ABI on a B200).
The reference for every check is plain torch fp32/fp64 on CPU (the oracle for single operators)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import gp as ogp


def to_padded_nhwc(x, pad=1):
    """[B,C,H,W] -> [B,H+2p,W+2p,C] with a zero border."""
    b, c, h, w = x.shape
    out = torch.zeros(b, h + 2 * pad, w + 2 * pad, c, dtype=x.dtype)
    out[:, pad:pad + h, pad:pad + w, :] = x.permute(0, 2, 3, 1)
    return out


def from_padded_nhwc(x, pad=1):
    b, hp, wp, c = x.shape
    return x[:, pad:hp - pad, pad:wp - pad, :].permute(0, 3, 1, 2).contiguous()


def _close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= atol + rtol * ref, "%s: max abs err %.3e vs ref scale %.3e" % (what, err, ref)


def check_conv1(lib, dev, B=3, H=10, W=37, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(64, generator=g)
    ref = F.conv2d(x, w, b, padding=1)
    T = lib.conv1_tiles(H, W)
    y = torch.empty(B, H, W, 64, device=dev)
    part = torch.zeros(B * T * 128, device=dev)
    lib.conv1_fwd(x.to(dev), w.to(dev), b.to(dev), y, part, B, H, W, 0)
    _close(y.cpu().permute(0, 3, 1, 2), ref, what="conv1 fwd")
    p = part.cpu().view(B, T, 2, 64).sum(1)
    _close(p[:, 0], ref.sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv1 sum")
    _close(p[:, 1], (ref * ref).sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv1 sumsq")
    # wgrad
    gy = torch.randn(B, 64, H, W, generator=g)
    xr = x.clone().requires_grad_(False)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    (F.conv2d(xr, wr, br, padding=1) * gy).sum().backward()
    dw = torch.empty(64, 3, 3, 3, device=dev)
    db = torch.empty(64, device=dev)
    scratch = torch.empty(lib.conv1_wgrad_nsplit() * 28 * 64, device=dev)
    lib.conv1_wgrad(x.to(dev), gy.permute(0, 2, 3, 1).contiguous().to(dev), dw, db, scratch, B, H, W, 0)
    _close(dw, wr.grad, rtol=1e-4, atol=1e-4, what="conv1 wgrad")
    _close(db, br.grad, rtol=1e-4, atol=1e-4, what="conv1 bgrad")


def check_conv3x3(lib, dev, B=3, H=6, W=5, seed=1, fwd_name="conv3x3_fwd", wgrad_name="conv3x3_wgrad", rtol=1e-4):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g)
    gy = torch.randn(B, 64, H, W, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=1)
    (ref * gy).sum().backward()
    wt_f = torch.empty(9, 64, 64, device=dev)
    wt_d = torch.empty(9, 64, 64, device=dev)
    lib.prep_weights(w.to(dev), wt_f, wt_d, 0)
    a = to_padded_nhwc(x).to(dev)
    T = lib.conv3x3_tiles(H, W)
    y = torch.full((B, H + 2, W + 2, 64), float("nan"), device=dev)
    part = torch.zeros(B * T * 128, device=dev)
    getattr(lib, fwd_name)(a, wt_f, b.to(dev), y, part, B, H, W, 0)
    _close(from_padded_nhwc(y.cpu()), ref, rtol=rtol, what="conv3x3 fwd")
    p = part.cpu().view(B, T, 2, 64).sum(1)
    _close(p[:, 0], ref.detach().sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv3x3 sum")
    _close(p[:, 1], (ref.detach() ** 2).sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv3x3 sumsq")
    # dgrad = same kernel on gy with flipped weights
    gyp = to_padded_nhwc(gy).to(dev)
    gx = torch.full((B, H + 2, W + 2, 64), float("nan"), device=dev)
    getattr(lib, fwd_name)(gyp, wt_d, None, gx, None, B, H, W, 0)
    _close(from_padded_nhwc(gx.cpu()), xr.grad, rtol=rtol, what="conv3x3 dgrad")
    # wgrad
    dw = torch.empty(64, 64, 3, 3, device=dev)
    db = torch.empty(64, device=dev)
    scratch = torch.empty(lib.conv3x3_wgrad_scratch_floats(), device=dev)
    getattr(lib, wgrad_name)(a, gyp, dw, db, scratch, B, H, W, 0)
    _close(dw, wr.grad, rtol=rtol, atol=1e-4, what="conv3x3 wgrad")
    _close(db, br.grad, rtol=rtol, atol=1e-4, what="conv3x3 bgrad")


def check_bn_relu_pool(lib, dev, E=2, ipe=3, H=7, W=6, pool=1, in_pad=1, out_pad=1, seed=2):
    g = torch.Generator().manual_seed(seed)
    B = E * ipe
    y = torch.randn(B, 64, H, W, generator=g) * 1.5 + 0.3
    gamma = torch.rand(64, generator=g) + 0.5
    gamma[3] = -0.7   # a negative scale must still pool correctly
    beta = torch.randn(64, generator=g) * 0.3
    rm0, rv0 = torch.randn(64, generator=g) * 0.1, torch.rand(64, generator=g) + 0.5
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    gout = torch.randn(B, 64, Ho, Wo, generator=g)
    # reference in float64: one BatchNorm batch per episode, sequential running-stat updates
    yr = y.double().clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    outs = []
    for e in range(E):
        o = F.batch_norm(yr[e * ipe:(e + 1) * ipe], rm, rv, gr, br, True, 0.1, 1e-5)
        o = F.relu(o)
        outs.append(F.max_pool2d(o, 2) if pool else o)
    ref = torch.cat(outs, 0)
    (ref * gout.double()).sum().backward()
    # device: partial sums as the conv epilogue would emit them (one "tile" per image)
    part = torch.stack([y.sum((2, 3)), (y * y).sum((2, 3))], 1).reshape(-1).to(dev)   # [B][1][2][64]
    mean = torch.empty(E, 64, device=dev)
    invstd = torch.empty(E, 64, device=dev)
    drm, drv = rm0.clone().to(dev), rv0.clone().to(dev)
    scratch_d = torch.empty(lib.bn_scratch_doubles(E), device=dev, dtype=torch.float64)
    lib.bn_finalize(part, B, 1, ipe, H * W, mean, invstd, drm, drv, scratch_d, 0.1, 1e-5, 0)
    _close(drm, rm, what="running_mean")
    _close(drv, rv, what="running_var")
    yd = (to_padded_nhwc(y) if in_pad else y.permute(0, 2, 3, 1).contiguous()).to(dev)
    out = torch.zeros(B, Ho + 2 * out_pad, Wo + 2 * out_pad, 64, device=dev)
    lib.bn_relu_pool_fwd(yd, mean, invstd, gamma.to(dev), beta.to(dev), out, B, H, W, ipe, in_pad, out_pad, pool, 0)
    got = from_padded_nhwc(out.cpu()) if out_pad else out.cpu().permute(0, 3, 1, 2)
    _close(got, ref, what="bn_relu_pool fwd")
    if out_pad:
        assert float(out.cpu()[:, 0].abs().max()) == 0.0
    # backward
    gd = (to_padded_nhwc(gout) if out_pad else gout.permute(0, 2, 3, 1).contiguous()).to(dev)
    gy = torch.zeros_like(yd)
    dg, db = torch.empty(64, device=dev), torch.empty(64, device=dev)
    chunks = lib.bn_bwd_chunks(H, W, pool)
    partial = torch.empty(B * chunks * 128, device=dev)
    sums = torch.empty(E * 128, device=dev)
    lib.bn_relu_pool_bwd(yd, gd, mean, invstd, gamma.to(dev), beta.to(dev), gy, dg, db, partial, sums, scratch_d, B, H,
                         W, ipe, in_pad, out_pad, pool, 0)
    got_gy = from_padded_nhwc(gy.cpu()) if in_pad else gy.cpu().permute(0, 3, 1, 2)
    _close(got_gy, yr.grad, rtol=1e-4, atol=1e-6, what="bn_relu_pool bwd gy")
    _close(dg, gr.grad, rtol=1e-4, atol=1e-5, what="dgamma")
    _close(db, br.grad, rtol=1e-4, atol=1e-5, what="dbeta")
    # eval mode
    em, ei = torch.empty(64, device=dev), torch.empty(64, device=dev)
    lib.bn_eval_prepare(drm, drv, em, ei, 64, 1e-5, 0)
    out2 = torch.zeros_like(out)
    lib.bn_relu_pool_fwd(yd, em, ei, gamma.to(dev), beta.to(dev), out2, B, H, W, 0, in_pad, out_pad, pool, 0)
    o = F.relu(F.batch_norm(y.double(), rm, rv, gamma.double(), beta.double(), False, 0.1, 1e-5))
    ref2 = F.max_pool2d(o, 2) if pool else o
    got2 = from_padded_nhwc(out2.cpu()) if out_pad else out2.cpu().permute(0, 3, 1, 2)
    _close(got2, ref2, what="bn_relu_pool eval")


def check_head(lib, dev, E=2, N=7, Cch=8, P=4, seed=3):
    g = torch.Generator().manual_seed(seed)
    D = Cch * P
    f_ref = torch.randn(E, N, D, generator=g) * 2 + 0.5           # reference (NCHW-flatten) feature order
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g) * 0.2
    rm0, rv0 = torch.randn(D, generator=g) * 0.1, torch.rand(D, generator=g) + 0.5
    gzh = torch.randn(E, N, D, generator=g)
    perm = torch.tensor([(j % Cch) * P + j // Cch for j in range(D)])   # nhwc index j -> reference index
    fr = f_ref.double().clone().requires_grad_(True)           # reference in float64
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    zs = []
    for e in range(E):
        z = F.batch_norm(fr[e], rm, rv, gr, br, True, 0.1, 1e-5)
        zs.append(F.normalize(z, p=2, dim=1))
    zh_ref = torch.stack(zs)
    (zh_ref * gzh.double()).sum().backward()
    f_dev = f_ref[:, :, perm].contiguous().to(dev)                      # our NHWC-flatten order
    z = torch.empty(E, N, D, device=dev)
    zh = torch.empty(E, N, D, device=dev)
    mean, invstd, var = (torch.empty(E, D, device=dev) for _ in range(3))
    inv = torch.empty(E * N, device=dev)
    drm, drv = rm0.clone().to(dev), rv0.clone().to(dev)
    lib.bn1d_fwd(f_dev, gamma.to(dev), beta.to(dev), drm, drv, z, mean, invstd, var, E, N, D, Cch, P, 1, 1, 0.1, 1e-5, 0)
    lib.l2norm_fwd(z, zh, inv, E * N, D, 1e-12, 0)
    _close(zh.cpu(), zh_ref[:, :, perm], what="head fwd")
    _close(drm, rm, what="bn_out running_mean")
    _close(drv, rv, what="bn_out running_var")
    gz = torch.empty(E, N, D, device=dev)
    gf = torch.empty(E, N, D, device=dev)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    pgrad = torch.empty(E * 2 * D, device=dev)
    lib.l2norm_bwd(zh, gzh[:, :, perm].contiguous().to(dev), inv, gz, E * N, D, 0)
    lib.bn1d_bwd(f_dev, gz, gamma.to(dev), mean, invstd, gf, dg, db, pgrad, E, N, D, Cch, P, 0)
    _close(gf.cpu(), fr.grad[:, :, perm], rtol=1e-4, atol=1e-6, what="head bwd")
    _close(dg, gr.grad, rtol=1e-4, atol=1e-5, what="bn_out dgamma")
    _close(db, br.grad, rtol=1e-4, atol=1e-5, what="bn_out dbeta")
    # eval mode
    z2 = torch.empty(E, N, D, device=dev)
    lib.bn1d_fwd(f_dev, gamma.to(dev), beta.to(dev), drm, drv, z2, None, None, None, E, N, D, Cch, P, 0, 0, 0.1, 1e-5, 0)
    ref2 = F.batch_norm(f_ref.double().view(E * N, D), rm, rv, gamma.double(), beta.double(), False, 0.1, 1e-5).view(E, N, D)
    _close(z2.cpu(), ref2[:, :, perm], what="bn_out eval")


def _envelope(ref32, ref64, k=4.0, tol=1e-4):
    """Gradient bar: `tol`, or k x the fp32 oracle's own max-norm distance to the fp64 oracle where that is larger (an
    ill-conditioned quantity cannot be reproduced better than the reference's own fp32 evaluation reproduces it)."""
    a, b = ref32.detach().double().reshape(-1), ref64.detach().double().reshape(-1)
    floor = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    return max(tol, k * floor)


def check_gp(lib, dev, E=3, C=3, per_class=4, D=24, M=9, seed=4, tol=1e-4, large=False):
    """Gram -> C exact-GP systems -> gradients -> prediction against the oracle in float64.  OUTPUTS (loss, predictive
    mean, L^-1) are held to `tol` = 1e-4 (north_star) at every size; GRADIENTS to max(1e-4, 4 x the fp32 oracle's own
    distance to float64) -- printed per quantity."""
    g = torch.Generator().manual_seed(seed)
    N = C * per_class
    z = F.normalize(torch.randn(E, N, D, generator=g), dim=2)
    zt = F.normalize(torch.randn(E, M, D, generator=g), dim=2)
    targets = -torch.ones(C, N)
    for c in range(C):
        targets[c, c * per_class:(c + 1) * per_class] = 1.0

    def oracle(dt):
        p = ogp.default_gp_params("bncossim", C, D, dtype=dt)
        p["raw_outputscale"] = torch.linspace(-0.4, 0.9, C, dtype=dt)
        p["constant"] = torch.linspace(-0.1, 0.2, C, dtype=dt)
        zr = z.to(dt).clone().requires_grad_(True)
        p["raw_outputscale"].requires_grad_(True)
        p["constant"].requires_grad_(True)
        losses = [ogp.mll_loss("bncossim", zr[e], targets.to(dt), p) for e in range(E)]
        (sum(losses) / E).backward()
        with torch.no_grad():
            pd = {k: v.detach() for k, v in p.items()}
            mean = torch.stack([ogp.predict("bncossim", z[e].to(dt), targets.to(dt), zt[e].to(dt), pd) for e in range(E)])
        return p, zr, torch.stack([l.detach() for l in losses]), mean
    p, zr, loss32, mean32 = oracle(torch.float32)
    p64, zr64, loss64, mean64 = oracle(torch.float64)
    zd = z.to(dev)
    gram = torch.empty(E, N, N, device=dev)
    lib.gram(zd, zd, gram, E, N, N, D, 0)
    _close(gram, z.double() @ z.double().transpose(1, 2), rtol=1e-5, atol=1e-6, what="gram")
    alpha = torch.empty(E, C, N, device=dev)
    lt = torch.empty(E, C, device=dev)
    info = torch.ones(E, C, device=dev, dtype=torch.int32)
    dk = torch.empty(E, C, N, N, device=dev)
    dh = torch.empty(E, C, 3, device=dev)
    ros, cst, rn = (p[k].detach().to(dev) for k in ("raw_outputscale", "constant", "raw_noise"))
    linv = torch.empty(E, C, N, N, device=dev)
    if large:       # the global-workspace blocked factorisation (csrc/gp_large.cu), same contract
        work = torch.empty(lib.gp_large_work_floats(E, C, N), device=dev)

        def fit(kb, a, li, l_, i_, dk_, dh_, gs):
            lib.gp_fit_large(kb, 0, targets.to(dev), 0, ros, cst, rn, a, li, l_, i_, dk_, dh_, work, gs, 0.0, E, C, N, 0)
    else:
        def fit(kb, a, li, l_, i_, dk_, dh_, gs):
            lib.gp_fit(kb, 0, targets.to(dev), 0, ros, cst, rn, a, li, l_, i_, dk_, dh_, gs, 0.0, E, C, N, 0)
    fit(gram, alpha, linv, lt, info, dk, dh, 1.0 / E)
    assert int(info.cpu().abs().sum()) == 0
    with torch.no_grad():       # L^-1 of K~ (lower triangular), used by the predictive variance
        s0 = F.softplus(p64["raw_outputscale"].detach()[0])
        kt = s0 * (z[0].double() @ z[0].double().T) + (F.softplus(p64["raw_noise"].detach()[0]) + 1e-4) * torch.eye(N, dtype=torch.float64)
        li_ref = torch.linalg.inv(torch.linalg.cholesky(kt))
    _close(linv[0, 0], li_ref, rtol=tol, atol=1e-6, what="L^-1")
    loss = torch.empty(E, device=dev)
    hyper = torch.empty(C, 3, device=dev)
    lib.gp_reduce(lt, dh, loss, hyper, E, C, 0)
    _close(loss, loss64, rtol=tol, atol=0.0, what="mll loss")
    bars = {"d raw_outputscale": _envelope(p["raw_outputscale"].grad, p64["raw_outputscale"].grad),
            "d constant": _envelope(p["constant"].grad, p64["constant"].grad), "d z": _envelope(zr.grad, zr64.grad)}
    _close(hyper[:, 0], p64["raw_outputscale"].grad, rtol=bars["d raw_outputscale"], atol=1e-7, what="d raw_outputscale")
    _close(hyper[:, 1], p64["constant"].grad, rtol=bars["d constant"], atol=1e-7, what="d constant")
    dz = torch.empty(E, N, D, device=dev)
    lib.gram_bwd(dk, zd, dz, E, C, N, D, 1.0, 0)
    _close(dz, zr64.grad, rtol=bars["d z"], atol=1e-7, what="d z")
    # prediction
    kx = torch.empty(E, M, N, device=dev)
    lib.gram(zt.to(dev), zd, kx, E, M, N, D, 0)
    mean = torch.empty(E, C, M, device=dev)
    pred = torch.empty(E, M, device=dev, dtype=torch.int32)
    lib.gp_predict(kx, 0, alpha, ros, cst, mean, pred, E, C, M, N, 0)
    _close(mean, mean64, rtol=tol, atol=1e-6, what="predictive mean")
    ref_pred = torch.sigmoid(mean32).numpy().argmax(axis=1)
    assert np.array_equal(pred.cpu().numpy(), ref_pred)
    print("check_gp N=%d%s: outputs at %.0e; gradient bars %s" % (N, " (tiled)" if large else "", tol,
                                                                   {k: "%.1e" % v for k, v in bars.items()}))
    # a non-PD system must be reported, not silently factorised
    bad = -torch.eye(N).repeat(E, 1, 1).to(dev)
    fit(bad, alpha, None, lt, info, None, None, 1.0)
    assert int((info.cpu() > 0).sum()) == E * C


def check_adam(lib, dev, n=1000, seed=5):
    g = torch.Generator().manual_seed(seed)
    p0 = torch.randn(n, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    pd, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * (10.0 ** (step - 3))
        pr.grad = gr.clone()
        opt.step()
        lib.adam_step(pd, gr.to(dev), m, v, n, 1e-3, 0.9, 0.999, 1e-8, step, 1.0, 0)
    _close(pd, pr.detach(), rtol=1e-6, atol=1e-7, what="adam")


def check_conv3x3_tc(lib, dev, B=3, H=6, W=5, seed=21, rtol=2e-5, fn="conv3x3_tc_fwd"):
    """tcgen05 3xTF32 forward / dgrad kernel against fp64 convolution (fp32-class accuracy expected)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g)
    gy = torch.randn(B, 64, H, W, generator=g)
    xr = x.double().requires_grad_(True)
    ref = F.conv2d(xr, w.double(), b.double(), padding=1)
    (ref * gy.double()).sum().backward()
    wb_f = torch.empty(lib.conv3x3_tc_weight_floats(), device=dev)
    wb_d = torch.empty(lib.conv3x3_tc_weight_floats(), device=dev)
    lib.prep_weights_tc(w.to(dev), wb_f, wb_d, 0)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    a = to_padded_nhwc(x).to(dev)
    T = lib.conv3x3_tiles(H, W)
    y = torch.full((B, H + 2, W + 2, 64), float("nan"), device=dev)
    part = torch.zeros(B * T * 128, device=dev)
    getattr(lib, fn)(a, wb_f, b.to(dev), y, part, err, B, H, W, 0)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "pipeline time-out"
    _close(from_padded_nhwc(y.cpu()), ref.detach(), rtol=rtol, atol=1e-6, what="tc conv fwd")
    p = part.cpu().view(B, T, 2, 64).sum(1)
    _close(p[:, 0], ref.detach().sum((2, 3)), rtol=1e-4, atol=1e-3, what="tc conv sum")
    _close(p[:, 1], (ref.detach() ** 2).sum((2, 3)), rtol=1e-4, atol=1e-3, what="tc conv sumsq")
    gx = torch.full((B, H + 2, W + 2, 64), float("nan"), device=dev)
    getattr(lib, fn)(to_padded_nhwc(gy).to(dev), wb_d, None, gx, None, err, B, H, W, 0)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "pipeline time-out"
    _close(from_padded_nhwc(gx.cpu()), xr.grad, rtol=rtol, atol=1e-6, what="tc conv dgrad")


def check_conv3x3_wgrad_tc(lib, dev, B=3, H=6, W=5, seed=31, rtol=2e-5):
    """tcgen05 wgrad (A = X^T staged in TMEM, B = gy transposed in smem) against fp64 autograd."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 64, H, W, generator=g)
    gy = torch.randn(B, 64, H, W, generator=g)
    wr = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).double().requires_grad_(True)
    br = torch.zeros(64, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x.double(), wr, br, padding=1) * gy.double()).sum().backward()
    dw = torch.full((64, 64, 3, 3), float("nan"), device=dev)
    db = torch.full((64,), float("nan"), device=dev)
    scratch = torch.empty(lib.conv3x3_wgrad_scratch_floats(), device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib.conv3x3_wgrad_tc(to_padded_nhwc(x).to(dev), to_padded_nhwc(gy).to(dev), dw, db, scratch, err, B, H, W, 0)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "pipeline time-out"
    _close(dw, wr.grad, rtol=rtol, atol=1e-5, what="tc wgrad")
    _close(db, br.grad, rtol=rtol, atol=1e-4, what="tc bgrad")


KIND = {"linear": 0, "rbf": 1, "matern": 2, "poli1": 3, "poli2": 4}
PARAM_NAME = {"linear": "raw_variance", "rbf": "raw_lengthscale", "matern": "raw_lengthscale", "poli1": "raw_offset",
              "poli2": "raw_offset"}


def check_gp_family(lib, dev, kernel, E=2, C=3, per_class=3, D=12, M=7, seed=40, tol=1e-4):
    """Kernel family (linear / rbf / matern / poli1 / poli2): centre -> Gram -> kernel epilogue -> gp_fit -> backward
    through the epilogue and the Gram, prediction with mean and variance -- against the oracle in float64.  OUTPUTS (loss,
    predictive mean, predictive variance) at `tol` = 1e-4; GRADIENTS at max(1e-4, 4 x fp32-oracle-vs-float64)."""
    g = torch.Generator().manual_seed(seed)
    N = C * per_class
    scale = 1.2 / D ** 0.5 if kernel in ("rbf", "matern") else 0.4     # keep ||x - x'||^2 / l^2 = O(1)
    z = torch.randn(E, N, D, generator=g) * scale
    zt = torch.randn(E, M, D, generator=g) * scale
    targets = -torch.ones(C, N)
    for c in range(C):
        targets[c, c * per_class:(c + 1) * per_class] = 1.0
    pn = PARAM_NAME[kernel]

    def oracle(dt):
        p = ogp.default_gp_params(kernel, C, D, dtype=dt)
        p["raw_outputscale"] = torch.linspace(-0.4, 0.5, C, dtype=dt)
        p["constant"] = torch.linspace(-0.1, 0.2, C, dtype=dt)
        p[pn] = torch.linspace(0.3, 1.2, C, dtype=dt)
        zr = z.to(dt).clone().requires_grad_(True)
        for k in ("raw_outputscale", "constant", pn):
            p[k].requires_grad_(True)
        losses = [ogp.mll_loss(kernel, zr[e], targets.to(dt), p) for e in range(E)]
        (sum(losses) / E).backward()
        with torch.no_grad():
            pd = {k: v.detach() for k, v in p.items()}
            ref = [ogp.predict(kernel, z[e].to(dt), targets.to(dt), zt[e].to(dt), pd, want_var=True) for e in range(E)]
        return p, zr, torch.stack([l.detach() for l in losses]), ref
    p, zr, loss32, ref32 = oracle(torch.float32)
    p64, zr64, loss64, ref64 = oracle(torch.float64)
    bar = lambda name, a, b: _envelope(a, b, tol=tol)
    kind = KIND[kernel]
    centred = kernel in ("rbf", "matern")
    zd = z.to(dev)
    xc = zd
    gram = torch.empty(E, N, N, device=dev)
    d2 = torch.empty(E, N, N, device=dev)
    lib.gram(xc, xc, gram, E, N, N, D, 0)
    lib.sqdist(xc, xc, d2, E, N, N, D, 0)
    _close(d2, torch.cdist(z.double(), z.double()) ** 2, rtol=1e-5, atol=1e-6, what="sqdist")
    assert float(d2.cpu().diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    sq = torch.empty(E, N, device=dev)
    lib.row_sqnorm(xc, sq, E * N, D, 0)
    rp = p[pn].detach().to(dev)
    kb = torch.empty(E, C, N, N, device=dev)
    lib.kernel_fwd(kind, None if centred else gram, d2 if centred else None, rp, kb, E, C, N, N, 0)
    alpha = torch.empty(E, C, N, device=dev)
    linv = torch.empty(E, C, N, N, device=dev)
    lt = torch.empty(E, C, device=dev)
    info = torch.ones(E, C, device=dev, dtype=torch.int32)
    dk = torch.empty(E, C, N, N, device=dev)
    dh = torch.empty(E, C, 3, device=dev)
    ros, cst, rn = (p[k].detach().to(dev) for k in ("raw_outputscale", "constant", "raw_noise"))
    lib.gp_fit(kb, N * N, targets.to(dev), 0, ros, cst, rn, alpha, linv, lt, info, dk, dh, 1.0 / E, 0.0, E, C, N, 0)
    assert int(info.cpu().abs().sum()) == 0
    loss = torch.empty(E, device=dev)
    hyper = torch.empty(C, 3, device=dev)
    lib.gp_reduce(lt, dh, loss, hyper, E, C, 0)
    _close(loss, loss64, rtol=tol, atol=0.0, what=kernel + " loss")
    _close(hyper[:, 0], p64["raw_outputscale"].grad, rtol=bar("s", p["raw_outputscale"].grad, p64["raw_outputscale"].grad),
           atol=1e-7, what=kernel + " d outputscale")
    _close(hyper[:, 1], p64["constant"].grad, rtol=bar("c", p["constant"].grad, p64["constant"].grad), atol=1e-7,
           what=kernel + " d constant")
    dg = torch.empty(E, N, N, device=dev)
    dparam = torch.empty(C, device=dev)
    scratch = torch.empty(E * C * N, device=dev)
    lib.kernel_bwd(kind, None if centred else gram, d2 if centred else None, rp, dk, dg, dparam, scratch, E, C, N, 0)
    _close(dparam, p64[pn].grad, rtol=bar("p", p[pn].grad, p64[pn].grad), atol=1e-7, what=kernel + " d " + pn)
    dz = torch.empty(E, N, D, device=dev)
    lib.gram_bwd(dg, xc, dz, E, 1, N, D, 1.0, 0)
    _close(dz, zr64.grad, rtol=bar("z", zr.grad, zr64.grad), atol=1e-7, what=kernel + " d z")
    # prediction: mean + variance
    ztd = zt.to(dev)
    xtc = ztd
    gx = torch.empty(E, M, N, device=dev)
    if centred:
        lib.sqdist(xtc, xc, gx, E, M, N, D, 0)
    else:
        lib.gram(xtc, xc, gx, E, M, N, D, 0)
    sqt = torch.empty(E, M, device=dev)
    lib.row_sqnorm(xtc, sqt, E * M, D, 0)
    kx = torch.empty(E, C, M, N, device=dev)
    lib.kernel_fwd(kind, None if centred else gx, gx if centred else None, rp, kx, E, C, M, N, 0)
    mean = torch.empty(E, C, M, device=dev)
    pred = torch.empty(E, M, device=dev, dtype=torch.int32)
    lib.gp_predict(kx, M * N, alpha, ros, cst, mean, pred, E, C, M, N, 0)
    # k(x*,x*) diagonal: kernel epilogue on the 1x1 "Gram" ||x*||^2
    kss = torch.empty(E, C, M, device=dev)
    for c in range(C):
        tmp = torch.empty(E * M, 1, 1, 1, device=dev)
        zeros = torch.zeros(E * M, 1, 1, device=dev)
        lib.kernel_fwd(kind, sqt.view(E * M, 1, 1).contiguous(), zeros, rp[c:c + 1].contiguous(), tmp, E * M, 1, 1, 1, 0)
        kss[:, c, :] = tmp.view(E, M)
    var = torch.empty(E, C, M, device=dev)
    lib.gp_predict_var(kx, M * N, kss.contiguous(), M, linv, ros, rn, var, E, C, M, N, 0)
    _close(mean, torch.stack([r[0] for r in ref64]), rtol=tol, atol=1e-6, what=kernel + " mean")
    _close(var, torch.stack([r[1] for r in ref64]), rtol=tol, atol=1e-6, what=kernel + " variance")


def check_conv2d(lib, dev, N=2, H=13, W=11, Cin=3, Cout=36, R=3, stride=2, pad=0, dil=2, relu=1, seed=50):
    """Generic NHWC convolution forward / dgrad / wgrad (+ fused ReLU) against torch autograd."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, R, R, generator=g) * 0.2
    b = torch.randn(Cout, generator=g) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, stride=stride, padding=pad, dilation=dil)
    if relu:
        ref = F.relu(ref)
    gy = torch.randn(ref.shape, generator=g)
    (ref * gy).sum().backward()
    Ho, Wo = ref.shape[2], ref.shape[3]
    assert lib.conv2d_out_size(H, R, stride, pad, dil) == Ho
    x_nhwc = torch.empty(N, H, W, Cin, device=dev)
    lib.nchw_to_nhwc(x.to(dev), x_nhwc, N, Cin, H, W, 0)
    out = torch.empty(N, Ho, Wo, Cout, device=dev)
    lib.conv2d_fwd(x_nhwc, w.to(dev), b.to(dev), out, N, H, W, Cin, Cout, R, R, stride, pad, dil, relu, 0)
    _close(out.cpu().permute(0, 3, 1, 2), ref.detach(), what="conv2d fwd")
    gyd = gy.permute(0, 2, 3, 1).contiguous().to(dev)
    gx = torch.empty(N, H, W, Cin, device=dev)
    lib.conv2d_dgrad(gyd, out, w.to(dev), gx, N, H, W, Cin, Cout, R, R, stride, pad, dil, relu, 0)
    _close(gx.cpu().permute(0, 3, 1, 2), xr.grad, rtol=1e-4, atol=1e-5, what="conv2d dgrad")
    dw = torch.empty(Cout, Cin, R, R, device=dev)
    db = torch.empty(Cout, device=dev)
    ns = lib.conv2d_wgrad_nsplit(N * Ho * Wo)
    scratch = torch.empty(ns * R * R * Cin * Cout, device=dev)
    lib.conv2d_wgrad(x_nhwc, gyd, out, dw, db, scratch, N, H, W, Cin, Cout, R, R, stride, pad, dil, relu, 0)
    _close(dw, wr.grad, rtol=1e-4, atol=1e-4, what="conv2d wgrad")
    _close(db, br.grad, rtol=1e-4, atol=1e-4, what="conv2d bgrad")
    if not relu and Cin < 16 and lib.conv2d_mma_ok(32, 32):      # narrow-input forward (the ResNet stem)
        wflat = torch.full((Cout, lib.conv2d_flat_k(Cin, R, R)), float("nan"), device=dev)
        lib.conv2d_prep_flat_mma(w.to(dev), wflat, Cout, Cin, R, R, 0)
        out3 = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev)
        lib.conv2d_fwd_flat_mma(x_nhwc, wflat, b.to(dev), out3, N, H, W, Cin, Cout, R, R, stride, pad, dil, 0)
        _close(out3.cpu().permute(0, 3, 1, 2), ref.detach(), rtol=2e-5, atol=2e-5, what="conv2d fwd (flat mma)")
    if not relu and lib.conv2d_mma_ok(Cin, Cout):       # tensor-core variants (ResNet layers)
        wf = torch.empty(R * R, Cout, Cin, device=dev)
        wd = torch.empty(R * R, Cin, Cout, device=dev)
        lib.conv2d_prep_mma(w.to(dev), wf, wd, Cout, Cin, R, R, 0)
        out2 = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev)
        lib.conv2d_fwd_mma(x_nhwc, wf, b.to(dev), out2, N, H, W, Cin, Cout, R, R, stride, pad, dil, 0)
        _close(out2.cpu().permute(0, 3, 1, 2), ref.detach(), rtol=2e-5, atol=2e-5, what="conv2d fwd (mma)")
        gx2 = torch.full((N, H, W, Cin), float("nan"), device=dev)
        lib.conv2d_dgrad_mma(gyd, wd, gx2, N, H, W, Cin, Cout, R, R, stride, pad, dil, 0)
        _close(gx2.cpu().permute(0, 3, 1, 2), xr.grad, rtol=1e-4, atol=2e-5, what="conv2d dgrad (mma)")


def check_spectral(lib, dev, E=2, N=6, M=5, Cch=4, P=3, Q=4, seed=60, rtol=3e-4):
    """Spectral-mixture kernel (ARD) forward / backward + GP fit and prediction against the oracle's autograd."""
    g = torch.Generator().manual_seed(seed)
    D = Cch * P
    x_ref = torch.randn(E, N, D, generator=g) * 0.3          # reference feature order
    xt_ref = torch.randn(E, M, D, generator=g) * 0.3
    y = torch.randn(E, N, generator=g)
    perm = torch.tensor([(j % Cch) * P + j // Cch for j in range(D)])
    p = ogp.default_gp_params("spectral", 1, D, classification=False)
    p["raw_mixture_weights"] = torch.randn(Q, generator=g) * 0.3
    p["raw_mixture_means"] = torch.randn(Q, 1, D, generator=g) * 0.5 - 1.0
    p["raw_mixture_scales"] = torch.randn(Q, 1, D, generator=g) * 0.5 - 0.5
    p["constant"] = torch.tensor([0.1])
    p["raw_noise"] = torch.tensor([-0.5])
    xr = x_ref.clone().requires_grad_(True)
    names = ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales", "constant", "raw_noise")
    for k in names:
        p[k].requires_grad_(True)
    losses = [ogp.mll_loss("spectral", xr[e], y[e:e + 1], p) for e in range(E)]
    (sum(losses) / E).backward()
    xd = x_ref[:, :, perm].contiguous().to(dev)                # our NHWC-flatten order
    rw, rmu, rv = (p[k].detach().reshape(Q, -1).contiguous().to(dev) if k != "raw_mixture_weights"
                   else p[k].detach().to(dev) for k in names[:3])
    kb = torch.empty(E, 1, N, N, device=dev)
    ec = torch.empty(E, Q, N, N, device=dev)
    lib.spectral_fwd(xd, xd, rw, rmu, rv, kb, ec, E, N, N, D, Q, Cch, P, 0)
    alpha = torch.empty(E, 1, N, device=dev)
    linv = torch.empty(E, 1, N, N, device=dev)
    lt = torch.empty(E, 1, device=dev)
    info = torch.ones(E, 1, device=dev, dtype=torch.int32)
    dk = torch.empty(E, 1, N, N, device=dev)
    dh = torch.empty(E, 1, 3, device=dev)
    cst, rn = p["constant"].detach().to(dev), p["raw_noise"].detach().to(dev)
    lib.gp_fit(kb, N * N, y.view(E, 1, N).contiguous().to(dev), N, None, cst, rn, alpha, linv, lt, info, dk, dh, 1.0 / E,
               0.0, E, 1, N, 0)
    assert int(info.cpu().abs().sum()) == 0
    loss = torch.empty(E, device=dev)
    hyper = torch.empty(1, 3, device=dev)
    lib.gp_reduce(lt, dh, loss, hyper, E, 1, 0)
    _close(loss, torch.stack([l.detach() for l in losses]), rtol=rtol, what="spectral loss")
    _close(hyper[:, 1], p["constant"].grad, rtol=rtol, atol=1e-6, what="spectral d constant")
    _close(hyper[:, 2], p["raw_noise"].grad, rtol=rtol, atol=1e-6, what="spectral d raw_noise")
    dw = torch.empty(Q, device=dev)
    dmu = torch.empty(Q, D, device=dev)
    dv = torch.empty(Q, D, device=dev)
    dx = torch.empty(E, N, D, device=dev)
    lib.spectral_bwd(xd, rw, rmu, rv, dk, ec, dw, dmu, dv, dx, E, N, D, Q, Cch, P, 0)
    _close(dw, p["raw_mixture_weights"].grad, rtol=rtol, atol=1e-6, what="d mixture weights")
    _close(dmu, p["raw_mixture_means"].grad.view(Q, D), rtol=rtol, atol=1e-6, what="d mixture means")
    _close(dv, p["raw_mixture_scales"].grad.view(Q, D), rtol=rtol, atol=1e-6, what="d mixture scales")
    _close(dx.cpu(), xr.grad[:, :, perm], rtol=rtol, atol=1e-6, what="spectral d x")
    # prediction
    xtd = xt_ref[:, :, perm].contiguous().to(dev)
    kx = torch.empty(E, 1, M, N, device=dev)
    lib.spectral_fwd(xtd, xd, rw, rmu, rv, kx, None, E, M, N, D, Q, Cch, P, 0)
    mean = torch.empty(E, 1, M, device=dev)
    lib.gp_predict(kx, M * N, alpha, None, cst, mean, None, E, 1, M, N, 0)
    kss = torch.nn.functional.softplus(p["raw_mixture_weights"].detach()).sum().expand(E, 1, M).contiguous().to(dev)
    var = torch.empty(E, 1, M, device=dev)
    lib.gp_predict_var(kx, M * N, kss, M, linv, None, rn, var, E, 1, M, N, 0)
    with torch.no_grad():
        pd = {k: v.detach() for k, v in p.items()}
        ref = [ogp.predict("spectral", x_ref[e], y[e:e + 1], xt_ref[e], pd, want_var=True) for e in range(E)]
    _close(mean, torch.stack([r[0] for r in ref]), rtol=rtol, atol=1e-5, what="spectral mean")
    _close(var, torch.stack([r[1] for r in ref]), rtol=rtol, atol=1e-5, what="spectral variance")


def check_conv1_tc(lib, dev, E=2, ipe=3, H=84, W=84, seed=70, rtol=2e-5):
    """tcgen05 first-layer convolution: mode 0 (y + partials), mode 1 (partials only), mode 2 (fused BN+ReLU+pool)."""
    g = torch.Generator().manual_seed(seed)
    B = E * ipe
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(64, generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    wb1 = torch.empty(2, 64, 32, device=dev)
    lib.prep_weights_conv1_tc(w.to(dev), wb1, 0)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    T = lib.conv1_tiles(H, W)
    y = torch.full((B, H, W, 64), float("nan"), device=dev)
    part = torch.zeros(B * T * 128, device=dev)
    lib.conv1_tc(x.to(dev), wb1, b.to(dev), y, part, None, None, None, None, None, err, B, H, W, ipe, 0, 0)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    _close(y.cpu().permute(0, 3, 1, 2), ref, rtol=rtol, atol=1e-6, what="conv1_tc y")
    p = part.cpu().view(B, T, 2, 64).sum(1)
    _close(p[:, 0], ref.sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv1_tc sum")
    _close(p[:, 1], (ref * ref).sum((2, 3)), rtol=1e-4, atol=1e-3, what="conv1_tc sumsq")
    part2 = torch.zeros(B * T * 128, device=dev)
    lib.conv1_tc(x.to(dev), wb1, b.to(dev), None, part2, None, None, None, None, None, err, B, H, W, ipe, 1, 0)
    assert torch.equal(part2, part)
    # fused apply with per-episode statistics
    gamma = torch.rand(64, generator=g) + 0.5
    gamma[5] = -0.6
    beta = torch.randn(64, generator=g) * 0.3
    mean = torch.randn(E, 64, generator=g) * 0.2
    invstd = torch.rand(E, 64, generator=g) + 0.5
    outs = []
    for e in range(E):
        r = ref[e * ipe:(e + 1) * ipe].float()
        z = (r - mean[e].view(1, -1, 1, 1)) * (invstd[e] * gamma).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        outs.append(F.max_pool2d(F.relu(z), 2))
    ref_act = torch.cat(outs, 0)
    act = torch.zeros(B, H // 2 + 2, W // 2 + 2, 64, device=dev)
    lib.conv1_tc(x.to(dev), wb1, b.to(dev), None, None, mean.to(dev), invstd.to(dev), gamma.to(dev), beta.to(dev), act, err,
                 B, H, W, ipe, 2, 0)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    _close(from_padded_nhwc(act.cpu()), ref_act, rtol=1e-5, atol=1e-5, what="conv1_tc fused apply")
    assert float(act.cpu()[:, 0].abs().max()) == 0.0 and float(act.cpu()[:, :, 0].abs().max()) == 0.0


def check_resnet_ops(lib, dev, E=2, ipe=2, H=9, W=7, C=8, seed=80):
    """Generic BatchNorm2d (+residual, +ReLU) fwd/bwd with per-episode statistics, MaxPool(3,2,1), global AvgPool."""
    g = torch.Generator().manual_seed(seed)
    B = E * ipe
    x = torch.randn(B, C, H, W, generator=g) * 1.3 + 0.2
    res = torch.randn(B, C, H, W, generator=g)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    rm0, rv0 = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    gy = torch.randn(B, C, H, W, generator=g)
    for relu, with_res in ((1, True), (0, False), (1, False)):
        xr, rr = x.double().clone().requires_grad_(True), res.double().clone().requires_grad_(True)     # float64 reference
        gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
        rm, rv = rm0.double().clone(), rv0.double().clone()
        outs = []
        for e in range(E):
            o = F.batch_norm(xr[e * ipe:(e + 1) * ipe], rm, rv, gr, br, True, 0.1, 1e-5)
            if with_res:
                o = o + rr[e * ipe:(e + 1) * ipe]
            outs.append(F.relu(o) if relu else o)
        ref = torch.cat(outs)
        (ref * gy.double()).sum().backward()
        nh = lambda t: t.permute(0, 2, 3, 1).contiguous().to(dev)
        xd, rd, gyd = nh(x), nh(res), nh(gy)
        mean, invstd = torch.empty(E, C, device=dev), torch.empty(E, C, device=dev)
        drm, drv = rm0.clone().to(dev), rv0.clone().to(dev)
        partial = torch.empty(lib.bn2d_partial_floats(B, H * W, C), device=dev)
        lib.bn2d_stats(xd, mean, invstd, drm, drv, partial, B, H * W, C, ipe, 0.1, 1e-5, 0)
        _close(drm, rm, what="bn2d running_mean")
        _close(drv, rv, what="bn2d running_var")
        y = torch.empty_like(xd)
        lib.bn2d_apply(xd, mean, invstd, gamma.to(dev), beta.to(dev), rd if with_res else None, y, B, H * W, C, ipe, relu, 0)
        _close(y.cpu().permute(0, 3, 1, 2), ref, what="bn2d fwd")
        gx, gres = torch.empty_like(xd), torch.empty_like(xd)
        dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
        sums = torch.empty(E * C * 2, device=dev)
        lib.bn2d_bwd(xd, y, gyd, mean, invstd, gamma.to(dev), gx, gres if with_res else None, dg, db, partial, sums, B,
                     H * W, C, ipe, relu, 0)
        _close(gx.cpu().permute(0, 3, 1, 2), xr.grad, rtol=1e-4, atol=1e-6, what="bn2d bwd gx")
        _close(dg, gr.grad, rtol=1e-4, atol=1e-5, what="bn2d dgamma")
        _close(db, br.grad, rtol=1e-4, atol=1e-5, what="bn2d dbeta")
        if with_res:
            _close(gres.cpu().permute(0, 3, 1, 2), rr.grad, rtol=1e-5, atol=1e-6, what="bn2d gres")
        if C % 4 == 0 and lib.has("dktb_bn2d_stats_l"):
            # layout-aware entry points: padded-flat buffers (border poisoned, must stay untouched) in every combination of
            # the tensor arguments give the dense results bit for bit
            def padded(t, fill=7.0):
                buf = torch.full((B, H + 2, W + 2, C), fill, device=dev)
                if t is not None:
                    buf[:, 1:-1, 1:-1] = t
                return buf

            def border_ok(buf, fill=7.0):
                bb = buf.clone()
                bb[:, 1:-1, 1:-1] = fill
                return bool((bb == fill).all())

            for lay in (31, 1, 2 | 8, 4 | 16, 1 | 4 | 8):
                lx, ly, lgy, lgx, lres = lay & 1, lay & 2, lay & 4, lay & 8, lay & 16
                x_a = padded(xd) if lx else xd
                r_a = (padded(rd) if lres else rd) if with_res else None
                mean2, invstd2 = torch.empty(E, C, device=dev), torch.empty(E, C, device=dev)
                drm2, drv2 = rm0.clone().to(dev), rv0.clone().to(dev)
                lib.bn2d_stats_l(x_a, mean2, invstd2, drm2, drv2, partial, B, H * W, C, ipe, 0.1, 1e-5, W, lay, 0)
                assert torch.equal(mean2, mean) and torch.equal(invstd2, invstd) and torch.equal(drm2, drm), "bn2d_stats_l"
                y_a = padded(None) if ly else torch.empty_like(xd)
                lib.bn2d_apply_l(x_a, mean, invstd, gamma.to(dev), beta.to(dev), r_a, y_a, B, H * W, C, ipe, relu, W, lay, 0)
                assert torch.equal(y_a[:, 1:-1, 1:-1] if ly else y_a, y), "bn2d_apply_l lay=%d" % lay
                assert not ly or border_ok(y_a)
                gy_a = padded(gyd) if lgy else gyd
                gx_a = padded(None) if lgx else torch.empty_like(xd)
                gres_a = (padded(None) if lres else torch.empty_like(xd)) if with_res else None
                dg2, db2 = torch.empty(C, device=dev), torch.empty(C, device=dev)
                lib.bn2d_bwd_l(x_a, y_a, gy_a, mean, invstd, gamma.to(dev), gx_a, gres_a, dg2, db2, partial, sums, B, H * W,
                               C, ipe, relu, W, lay, 0)
                assert torch.equal(gx_a[:, 1:-1, 1:-1] if lgx else gx_a, gx), "bn2d_bwd_l gx lay=%d" % lay
                assert torch.equal(dg2, dg) and torch.equal(db2, db), "bn2d_bwd_l parameter gradients"
                assert not lgx or border_ok(gx_a)
                if with_res:
                    assert torch.equal(gres_a[:, 1:-1, 1:-1] if lres else gres_a, gres), "bn2d_bwd_l gres"
                    assert not lres or border_ok(gres_a)
                acc = padded(xd) if lx else xd.clone()
                lib.add_inplace_l(acc, padded(gyd) if ly else gyd, B, H * W, C, W, lay & 3, 0)
                assert torch.equal(acc[:, 1:-1, 1:-1] if lx else acc, xd + gyd) and (not lx or border_ok(acc)), "add_inplace_l"
    # eval mode (ipe = 0: one statistics row)
    em, ei = torch.empty(C, device=dev), torch.empty(C, device=dev)
    lib.bn_eval_prepare(rm0.to(dev), rv0.to(dev), em, ei, C, 1e-5, 0)
    y2 = torch.empty(B, H, W, C, device=dev)
    lib.bn2d_apply(x.permute(0, 2, 3, 1).contiguous().to(dev), em, ei, gamma.to(dev), beta.to(dev), None, y2, B, H * W, C, 0, 1, 0)
    _close(y2.cpu().permute(0, 3, 1, 2), F.relu(F.batch_norm(x, rm0, rv0, gamma, beta, False, 0.1, 1e-5)), what="bn2d eval")
    # max pool 3x3 s2 p1
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    gp = torch.randn(ref.shape, generator=g)
    (ref * gp).sum().backward()
    Ho, Wo = ref.shape[2], ref.shape[3]
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    yp = torch.empty(B, Ho, Wo, C, device=dev)
    idx = torch.empty(B, Ho, Wo, C, device=dev, dtype=torch.uint8)
    lib.maxpool3_fwd(xd, yp, idx, B, H, W, C, 0)
    _close(yp.cpu().permute(0, 3, 1, 2), ref, what="maxpool fwd")
    gxp = torch.empty_like(xd)
    lib.maxpool3_bwd(gp.permute(0, 2, 3, 1).contiguous().to(dev), idx, gxp, B, H, W, C, 0)
    _close(gxp.cpu().permute(0, 3, 1, 2), xr.grad, what="maxpool bwd")
    # global average pool
    ya = torch.empty(B, C, device=dev)
    lib.avgpool_fwd(xd, ya, B, H * W, C, 0)
    _close(ya, x.mean((2, 3)), what="avgpool fwd")
    ga = torch.randn(B, C, generator=g)
    gxa = torch.empty_like(xd)
    lib.avgpool_bwd(ga.to(dev), gxa, B, H * W, C, 0)
    _close(gxa.cpu().permute(0, 3, 1, 2), (ga / (H * W)).view(B, C, 1, 1).expand(B, C, H, W), what="avgpool bwd")
    a, b2 = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    ad = a.clone().to(dev)
    lib.add_inplace(ad, b2.to(dev), 1000, 0)
    _close(ad, a + b2, what="add")


def check_conv1_bwd_fused(lib, dev, E=2, ipe=2, H=10, W=37, out_pad=1, seed=21, fn="conv1_bwd_fused"):
    """First block backward, fused (BN/ReLU/MaxPool backward inside the conv1 weight gradient), against torch autograd of
    conv1 -> per-episode BatchNorm -> ReLU -> MaxPool2d(2) and against the split device path."""
    g = torch.Generator().manual_seed(seed)
    B = E * ipe
    x = torch.randn(B, 3, H, W, generator=g)
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).double().requires_grad_(True)        # float64 reference
    b = torch.randn(64, generator=g).double().requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(64, generator=g)).double().requires_grad_(True)
    beta = (0.1 * torch.randn(64, generator=g)).double().requires_grad_(True)
    Ho, Wo = H // 2, W // 2
    gout = torch.randn(B, 64, Ho, Wo, generator=g)
    y = F.conv2d(x.double(), w, b, padding=1)
    outs = []
    for e in range(E):
        ye = y[e * ipe:(e + 1) * ipe]
        outs.append(F.max_pool2d(F.relu(F.batch_norm(ye, None, None, gamma, beta, True, 0.0, 1e-5)), 2))
    (torch.cat(outs) * gout.double()).sum().backward()
    yd = y.detach().float().permute(0, 2, 3, 1).contiguous().to(dev)
    yv = y.detach().view(E, ipe, 64, H, W)
    mean = yv.mean((1, 3, 4)).float().to(dev).contiguous()
    invstd = (1.0 / torch.sqrt(yv.var((1, 3, 4), unbiased=False) + 1e-5)).float().to(dev).contiguous()
    gpad = torch.zeros(B, Ho + 2 * out_pad, Wo + 2 * out_pad, 64)
    gpad[:, out_pad:out_pad + Ho, out_pad:out_pad + Wo] = gout.permute(0, 2, 3, 1)
    gpad = gpad.to(dev)
    chunks = lib.bn_bwd_chunks(H, W, 1)
    partial = torch.empty(B * chunks * 128, device=dev)
    sums = torch.empty(E * 128, device=dev)
    scratch_d = torch.empty(lib.bn_scratch_doubles(E), device=dev, dtype=torch.float64)
    dg, dbt = torch.empty(64, device=dev), torch.empty(64, device=dev)
    gd, bd = gamma.detach().float().to(dev), beta.detach().float().to(dev)
    lib.bn_relu_pool_bwd(yd, gpad, mean, invstd, gd, bd, None, dg, dbt, partial, sums, scratch_d, B, H, W, ipe, 0, out_pad, 1, 0)
    _close(dg, gamma.grad, rtol=1e-4, atol=1e-5, what="fused L0: d gamma")
    _close(dbt, beta.grad, rtol=1e-4, atol=1e-5, what="fused L0: d beta")
    dw, db = torch.empty(64, 3, 3, 3, device=dev), torch.empty(64, device=dev)
    scratch = torch.empty(lib.conv1_wgrad_nsplit() * 28 * 64, device=dev)
    getattr(lib, fn)(x.to(dev), yd, gpad, mean, invstd, gd, bd, sums, dw, db, scratch, B, H, W, ipe, out_pad, 0)
    _close(dw, w.grad, rtol=1e-4, atol=2e-5, what="fused L0: d conv1 weight")
    assert float(db.abs().max()) <= 1e-3 * float(w.grad.abs().max()) + 1e-5      # cancelled by BatchNorm
    # split path: same numbers up to summation order
    gy = torch.empty(B, H, W, 64, device=dev)
    lib.bn_relu_pool_bwd(yd, gpad, mean, invstd, gd, bd, gy, dg, dbt, partial, sums, scratch_d, B, H, W, ipe, 0, out_pad, 1, 0)
    dw2, db2 = torch.empty_like(dw), torch.empty_like(db)
    lib.conv1_wgrad(x.to(dev), gy, dw2, db2, scratch, B, H, W, 0)
    _close(dw, dw2, rtol=1e-4, atol=1e-4, what="fused vs split")


# ---------------------------------------------------------------------------------------------- episode feeder (8f-1)
def synth_image(seed, h, w):
    """Smooth structure + noise over the full 0..255 range (same generator as tests/golden/make_golden_transforms.py)."""
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = 127 + 90 * np.sin(xx / (7.0 + 3 * c) + seed) * np.cos(yy / (11.0 - 2 * c)) + r.randn(h, w) * 40
    return np.clip(img, 0, 255).astype(np.uint8)


def feeder_oracle(store_images, drawn, S, aug):
    """oracle/transforms.py applied image by image to a draw() result -> [n,3,S,S] float32."""
    from oracle import transforms as ot
    outs = []
    for p, f in zip(drawn["params"], drawn["factors"]):
        img = store_images[int(p[0])]
        if aug:
            outs.append(ot.transform_aug(img, (int(p[1]), int(p[2]), int(p[3]), int(p[4])), f if p[6] else None,
                                         int(p[5]), S))
        else:
            outs.append(ot.transform_plain(img, S))
    return np.stack(outs)


def check_episode_transform(lib, dev, S=12, shapes=((20, 31), (12, 12), (9, 14), (40, 26), (33, 50), (16, 11)),
                            n_way=2, batch=3, E=2, seed=90, tmp_budget=None):
    """Device feeder vs the PIL-pinned oracle, BIT-EXACT, aug and plain pipelines."""
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore, EpisodeFeeder
    images, labels = [], []
    for k, (h, w) in enumerate(shapes):
        for j in range(batch + 1):
            images.append(synth_image(seed + 10 * k + j, h + j, w + 2 * j))
            labels.append(k)
    store = EpisodeStore(images, labels, dev)
    for aug in (True, False):
        feed = EpisodeFeeder(store, S, n_way, batch - 1, 1, n_episode=E, aug=aug, seed=seed, lib=lib)
        if tmp_budget is not None:
            feed.SMEM_BUDGET = tmp_budget                       # force several bands per image
        drawn = feed.draw(E)
        if aug:                                                 # exercise the blend short-cuts and the clipping branch
            drawn["factors"][0] = (1.0, 0.0, 1.4)
            drawn["factors"][1] = (0.0, 1.0, 0.6)
            drawn["params"][2, 6] = 0
        got = feed.transform(drawn).cpu().numpy()
        feed.check()
        ref = feeder_oracle(images, drawn, S, aug)
        bad = int((got != ref).sum())
        assert bad == 0, "episode transform (aug=%s): %d of %d values differ, max |d| = %g" % (
            aug, bad, ref.size, float(np.abs(got - ref).max()))
        # labels / composition: class-major episodes, S+Q distinct members of each sampled class
        ids, lab = drawn["ids"], drawn["labels"]
        assert ids.shape == (E, n_way, batch)
        for e in range(E):
            assert len(set(lab[e, :, 0].tolist())) == n_way
            for j in range(n_way):
                assert len(set(ids[e, j].tolist())) == batch
                assert all(labels[i] == lab[e, j, 0] for i in ids[e, j])
    return store


def check_episode_transform_errors(lib, dev):
    """The kernel's error flag: a crop box outside its image, more taps than kmax."""
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore, EpisodeFeeder
    from deep_kernel_transfer_b200._lib import DktbError
    images = [synth_image(3 + i, 30, 40) for i in range(4)]
    store = EpisodeStore(images, [0, 0, 1, 1], dev)
    feed = EpisodeFeeder(store, 8, 2, 1, 1, n_episode=1, aug=True, seed=1, lib=lib)
    d = feed.draw(1)
    d["params"][1, 1:5] = (25, 0, 10, 10)                       # rows 25..34 of a 30-row image
    feed.transform(d)
    try:
        feed.check()
    except DktbError as ex:
        assert "code 3" in str(ex)
    else:
        raise AssertionError("out-of-image crop box was not reported")


def check_gp_jitter(lib, dev, N=12, large=False, seed=90):
    """psd_safe_cholesky semantics (GPyTorch utils/cholesky.py; the reference relies on them, README.md:27): a system whose
    K~ has a slightly negative eigenvalue factorises after the 2nd retry (1e-6 fails, 1e-5 passes) and its outputs are
    those of K~ + 1e-5 I; a plainly indefinite one fails every retry -> positive info, NaN loss, ZERO gradients; with
    jitter = 0 there is a single attempt."""
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(N, N, generator=g, dtype=torch.float64))
    noise = float(F.softplus(torch.tensor(-1.0)) + 1e-4)
    s = float(F.softplus(torch.tensor(0.3)))

    def system(min_eig):
        ev = torch.linspace(0.5, 2.0, N, dtype=torch.float64)
        ev[0] = min_eig
        kt = (q * ev) @ q.T
        return ((kt - noise * torch.eye(N, dtype=torch.float64)) / s).float(), kt
    E, C = 3, 1
    kb_ok, _ = system(0.3)
    kb_jit, kt_jit = system(-5e-6)
    kb_bad, _ = system(-1.0)
    kb = torch.stack([kb_ok, kb_jit, kb_bad]).view(E, 1, N, N).contiguous().to(dev)
    y = torch.randn(E, C, N, generator=g).to(dev)
    ros, cst, rn = (torch.tensor([v], device=dev) for v in (0.3, 0.1, -1.0))
    alpha = torch.full((E, C, N), 7.0, device=dev)
    lt = torch.zeros(E, C, device=dev)
    info = torch.zeros(E, C, device=dev, dtype=torch.int32)
    dk = torch.full((E, C, N, N), 7.0, device=dev)
    dh = torch.full((E, C, 3), 7.0, device=dev)

    def fit(jitter):
        if large:
            work = torch.empty(lib.gp_large_work_floats(E, C, N), device=dev)
            lib.gp_fit_large(kb, N * N, y, C * N, ros, cst, rn, alpha, None, lt, info, dk, dh, work, 1.0, jitter, E, C, N, 0)
        else:
            lib.gp_fit(kb, N * N, y, C * N, ros, cst, rn, alpha, None, lt, info, dk, dh, 1.0, jitter, E, C, N, 0)
    fit(1e-6)
    got = info.cpu().view(-1).tolist()
    assert got[0] == 0 and got[1] == -2 and got[2] > 0, got
    lt_c = lt.cpu().view(-1)
    assert bool(torch.isnan(lt_c[2])) and bool(torch.isfinite(lt_c[:2]).all())
    assert float(dk[2].abs().max()) == 0.0 and float(dh[2].abs().max()) == 0.0 and float(alpha[2].abs().max()) == 0.0
    # the retried system equals the plain factorisation of K~ + 1e-5 I
    ktj = kt_jit + 1e-5 * torch.eye(N, dtype=torch.float64)
    r = (y[1, 0].cpu().double() - 0.1)
    l = torch.linalg.cholesky(ktj)
    a_ref = torch.cholesky_solve(r.unsqueeze(-1), l).squeeze(-1)
    logp = -0.5 * (float(r @ a_ref) + 2.0 * float(torch.log(torch.diagonal(l)).sum()) + N * math.log(2 * math.pi))
    # the smallest eigenvalue of the retried matrix is 5e-6 +- fp32 rounding of its O(1) entries (~1e-7 .. 1e-6), and the
    # quadratic form is dominated by 1 / that eigenvalue: the comparison can only hold to ~20 %
    assert abs(float(lt_c[1]) - (-logp / N)) <= 0.25 * abs(logp / N), (float(lt_c[1]), -logp / N)
    # sticky status over several fits + a single attempt when jitter == 0
    sticky = torch.zeros(2, device=dev, dtype=torch.int32)
    lib.gp_info_accumulate(info, sticky, E * C, 0)
    fit(0.0)
    got0 = info.cpu().view(-1).tolist()
    assert got0[0] == 0 and got0[1] > 0 and got0[2] > 0, got0
    lib.gp_info_accumulate(info, sticky, E * C, 0)
    hi, lo = sticky.cpu().tolist()
    assert hi == max(got + got0) and lo == -2


def check_conv_tcg(lib, dev, B=2, H=6, W=5, Cin=64, Cout=128, R=3, seed=100, rtol=2e-5, bias=True):
    """ResNet-layer convolution on tcgen05 (csrc/conv_tcg.cu): forward (+bias) and dgrad against torch in float64;
    3x3 over padded-flat tensors (border of the output untouched), 1x1 as a GEMM over dense rows."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, R, R, generator=g) * (1.0 / (Cin * R * R) ** 0.5)
    b = torch.randn(Cout, generator=g) if bias else None
    xr = x.double().clone().requires_grad_(True)
    wr = w.double().clone().requires_grad_(True)
    br = b.double().clone().requires_grad_(True) if bias else None
    ref = F.conv2d(xr, wr, br, padding=R // 2)
    gout = torch.randn(ref.shape, generator=g)
    (ref * gout.double()).sum().backward()
    n = lib.conv_tcg_weight_floats(Cin, Cout, R)
    dw = torch.empty(Cout, Cin, R, R, device=dev)
    db = torch.empty(Cout, device=dev) if bias else None
    scratch = torch.empty(lib.wgrad_tcg_scratch_floats(B, H, W, Cin, Cout, R), device=dev)
    wf, wd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    lib.prep_weights_tcg(w.to(dev), wf, wd, Cout, Cin, R, 0)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    assert lib.conv_tcg_ok(Cin, Cout, R, 1, R // 2, 1, W)
    if R == 3:
        xp = to_padded_nhwc(x).to(dev)
        yp = torch.full((B, H + 2, W + 2, Cout), 7.0, device=dev)
        lib.conv_tcg(xp, wf, b.to(dev) if bias else None, yp, err, B, H, W, Cin, Cout, R, 0)
        assert int(err) == 0
        got = from_padded_nhwc(yp.cpu())
        assert float((yp.cpu()[:, 0] - 7.0).abs().max()) == 0.0 and float((yp.cpu()[:, :, 0] - 7.0).abs().max()) == 0.0
        gp = to_padded_nhwc(gout).to(dev)
        gxp = torch.zeros(B, H + 2, W + 2, Cin, device=dev)
        lib.conv_tcg(gp, wd, None, gxp, err, B, H, W, Cout, Cin, R, 0)
        gx = from_padded_nhwc(gxp.cpu())
        lib.wgrad_tcg(xp, gp, dw, db, scratch, err, B, H, W, Cin, Cout, R, 0)
        # pad_copy / zero_border round trip
        dense = x.permute(0, 2, 3, 1).contiguous().to(dev)
        pp = torch.zeros(B, H + 2, W + 2, Cin, device=dev)
        lib.pad_copy(dense, pp, B, H, W, Cin, 0, 0)
        assert torch.equal(pp, xp)
        pp.add_(1.0)
        lib.zero_border(pp, B, H, W, Cin, 0)
        assert torch.equal(pp[:, 1:-1, 1:-1], xp[:, 1:-1, 1:-1] + 1.0) and float(pp[:, 0].abs().max()) == 0.0 \
            and float(pp[:, -1].abs().max()) == 0.0 and float(pp[:, :, 0].abs().max()) == 0.0 and float(pp[:, :, -1].abs().max()) == 0.0
        back = torch.empty_like(dense)
        lib.pad_copy(back, xp, B, H, W, Cin, 1, 0)
        assert torch.equal(back, dense)
    else:
        xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
        yd = torch.empty(B, H, W, Cout, device=dev)
        lib.conv_tcg(xd, wf, b.to(dev) if bias else None, yd, err, B, H, W, Cin, Cout, R, 0)
        assert int(err) == 0
        got = yd.cpu().permute(0, 3, 1, 2)
        gd = gout.permute(0, 2, 3, 1).contiguous().to(dev)
        gxd = torch.empty(B, H, W, Cin, device=dev)
        lib.conv_tcg(gd, wd, None, gxd, err, B, H, W, Cout, Cin, R, 0)
        gx = gxd.cpu().permute(0, 3, 1, 2)
        lib.wgrad_tcg(xd, gd, dw, db, scratch, err, B, H, W, Cin, Cout, R, 0)
    assert int(err) == 0
    _close(got, ref.detach(), rtol=rtol, atol=1e-6, what="conv_tcg fwd %dx%d %d->%d" % (R, R, Cin, Cout))
    _close(gx, xr.grad, rtol=rtol, atol=1e-6, what="conv_tcg dgrad %dx%d %d->%d" % (R, R, Cin, Cout))
    _close(dw, wr.grad, rtol=rtol, atol=1e-5, what="conv_tcg wgrad %dx%d %d->%d" % (R, R, Cin, Cout))
    if bias:
        _close(db, br.grad, rtol=rtol, atol=1e-4, what="conv_tcg bias grad")


def check_conv_tcg_s2(lib, dev, B=2, H=8, W=10, C=64, Cout=128, seed=130, rtol=2e-5, bias=False, x_pad=False):
    """Stride-2 3x3 / pad 1 convolution through the space-to-depth path of csrc/conv_tcg.cu (dktb_s2d +
    dktb_conv_tcg_s2): forward and dgrad against torch in float64; pack / unpack round trip."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) * (1.0 / (C * 9) ** 0.5)
    b = torch.randn(Cout, generator=g) if bias else None
    xr = x.double().clone().requires_grad_(True)
    wr = w.double().clone().requires_grad_(True)
    br = b.double().clone().requires_grad_(True) if bias else None
    ref = F.conv2d(xr, wr, br, stride=2, padding=1)
    gout = torch.randn(ref.shape, generator=g)
    (ref * gout.double()).sum().backward()
    Ho, Wo = H // 2, W // 2
    assert lib.conv_tcg_s2_ok(C, Cout, H, W)
    n = lib.conv_tcg_s2_weight_floats(C, Cout)
    wf, wd = torch.full((n,), float("nan"), device=dev), torch.full((n,), float("nan"), device=dev)   # unused slots: never read
    lib.prep_weights_tcg_s2(w.to(dev), wf, wd, Cout, C, 0)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    xd = (to_padded_nhwc(x) if x_pad else x.permute(0, 2, 3, 1).contiguous()).to(dev)
    xs = torch.zeros(B, Ho + 2, Wo + 2, 4 * C, device=dev)
    lib.s2d(xd, xs, B, H, W, C, int(x_pad), 0, 0)
    yp = torch.full((B, Ho + 2, Wo + 2, Cout), 7.0, device=dev)
    lib.conv_tcg_s2(xs, wf, b.to(dev) if bias else None, yp, err, B, Ho, Wo, C, Cout, 0, 0)
    assert int(err) == 0
    _close(from_padded_nhwc(yp.cpu()), ref.detach(), rtol=rtol, atol=1e-6, what="conv_tcg_s2 fwd %d->%d" % (C, Cout))
    assert float((yp.cpu()[:, 0] - 7.0).abs().max()) == 0.0 and float((yp.cpu()[:, :, -1] - 7.0).abs().max()) == 0.0
    gp = to_padded_nhwc(gout).to(dev)
    dxs = torch.zeros(B, Ho + 2, Wo + 2, 4 * C, device=dev)
    lib.conv_tcg_s2(gp, wd, None, dxs, err, B, Ho, Wo, C, Cout, 1, 0)
    assert int(err) == 0
    gx = torch.full_like(xd, 7.0)
    lib.s2d(gx, dxs, B, H, W, C, int(x_pad), 1, 0)
    got = from_padded_nhwc(gx.cpu()) if x_pad else gx.cpu().permute(0, 3, 1, 2)
    _close(got, xr.grad, rtol=rtol, atol=1e-6, what="conv_tcg_s2 dgrad %d->%d" % (C, Cout))
    if x_pad:
        assert float((gx.cpu()[:, 0] - 7.0).abs().max()) == 0.0
    if lib.wgrad_tcg_s2_ok(C, Cout, H, W):
        dw = torch.full((Cout, C, 3, 3), 7.0, device=dev)
        db = torch.empty(Cout, device=dev) if bias else None
        scratch = torch.empty(lib.wgrad_tcg_s2_scratch_floats(B, Ho, Wo, C, Cout), device=dev)
        lib.wgrad_tcg_s2(xs, gp, dw, db, scratch, err, B, Ho, Wo, C, Cout, 0)
        assert int(err) == 0
        _close(dw, wr.grad, rtol=rtol, atol=1e-5, what="wgrad_tcg_s2 %d->%d" % (C, Cout))
        if bias:
            _close(db, br.grad, rtol=rtol, atol=1e-4, what="wgrad_tcg_s2 bias grad")
    else:
        assert W // 2 > 29
    # stride-2 1x1 shortcut helpers: gather of every second pixel, scatter with zeros elsewhere
    xg = torch.empty(B, Ho, Wo, C, device=dev)
    lib.subsample2(xd, xg, B, H, W, C, int(x_pad), 0, 0)
    x_nhwc = x.permute(0, 2, 3, 1)
    assert torch.equal(xg.cpu(), x_nhwc[:, ::2, ::2, :].contiguous()), "subsample2 gather"
    sc = torch.full_like(xd, 7.0)
    lib.subsample2(sc, xg, B, H, W, C, int(x_pad), 1, 0)
    want = torch.zeros(B, H, W, C)
    want[:, ::2, ::2, :] = x_nhwc[:, ::2, ::2, :]
    assert torch.equal(sc.cpu()[:, 1:-1, 1:-1] if x_pad else sc.cpu(), want), "subsample2 scatter"
    back = torch.zeros_like(xd)
    lib.s2d(back, xs, B, H, W, C, int(x_pad), 1, 0)
    assert torch.equal(back[:, 1:-1, 1:-1] if x_pad else back, xd[:, 1:-1, 1:-1] if x_pad else xd), "s2d round trip"


def check_stem_tc(lib, dev, B=2, H=32, W=32, seed=120, rtol=2e-5, bias=False):
    """ResNet stem (7x7, stride 2, pad 3, 3 -> 64) on tcgen05 (csrc/stem_tc.cu) against torch in float64."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * (1.0 / 147 ** 0.5)
    b = torch.randn(64, generator=g) if bias else None
    wr = w.double().clone().requires_grad_(True)
    ref = F.conv2d(x.double(), wr, b.double() if bias else None, stride=2, padding=3)
    gout = torch.randn(ref.shape, generator=g)
    (ref * gout.double()).sum().backward()
    ref = ref.detach()
    assert lib.stem_tc_ok(3, 64, 7, 2, 3, 1, H, W)
    wb = torch.empty(lib.stem_tc_weight_floats(), device=dev)
    lib.prep_weights_stem_tc(w.to(dev), wb, 0)
    y = torch.full((B, H // 2, W // 2, 64), 7.0, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    lib.stem_tc(x.to(dev), wb, b.to(dev) if bias else None, y, err, B, H, W, 0)
    assert int(err) == 0
    _close(y.cpu().permute(0, 3, 1, 2), ref, rtol=rtol, atol=1e-6, what="stem_tc %dx%d" % (H, W))
    dw = torch.full((64, 3, 7, 7), 7.0, device=dev)
    scratch = torch.empty(lib.stem_wgrad_tc_scratch_floats(B, H), device=dev)
    lib.stem_wgrad_tc(x.to(dev), gout.permute(0, 2, 3, 1).contiguous().to(dev), dw, scratch, err, B, H, W, 0)
    assert int(err) == 0
    _close(dw, wr.grad, rtol=rtol, atol=1e-5, what="stem_wgrad_tc %dx%d" % (H, W))


def check_gram_tc(lib, dev, E=3, M=105, N=105, D=1600, seed=110, same=True):
    """tcgen05 Gram / cross-kernel product against float64 (3xTF32: fp32-class accuracy), incl. ragged tiles."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(E, M, D, generator=g)
    x2 = x1 if same else torch.randn(E, N, D, generator=g)
    assert lib.gram_tc_ok(E, M, N, D)
    a = x1.to(dev)
    b = a if same else x2.to(dev)
    out = torch.full((E, M, N), 7.0, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    scratch = torch.empty(lib.gram_tc_scratch_floats(E, M, N, D), device=dev)
    lib.gram_tc(a, b, out, scratch, err, E, M, N, D, 0)
    assert int(err) == 0
    ref = x1.double() @ x2.double().transpose(1, 2)
    _close(out, ref, rtol=2e-6, atol=1e-6 * D ** 0.5, what="gram_tc %dx%dx%d" % (M, N, D))
    if same:
        o = out.cpu()
        assert float((o - o.transpose(1, 2)).abs().max()) <= 1e-5 * float(o.abs().max())
