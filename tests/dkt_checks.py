"""End-to-end parity checks of the drop-in ``DKT`` module against the CPU oracle (oracle/episode.py),
shared by the emulation tests (CPU, host logic) and the GPU tests (parity proper)."""
import numpy as np
import torch

from oracle import episode as oep


def load_oracle_params(model, oracle):
    """Copy the oracle's backbone + GP parameters into the module (reference state_dict key names)."""
    sd = model.feature.state_dict()
    new = {}
    for k in sd:
        kk = k
        if kk not in oracle.bb:                     # ConvBlock aliases: trunk.i.trunk.{0,1}.*
            head, tail = kk.split(".trunk.", 1)
            idx, rest = tail.split(".", 1)
            kk = head + (".C." if idx == "0" else ".BN.") + rest
        new[k] = oracle.bb[kk].detach().clone()
    model.feature.load_state_dict(new)
    for c, m in enumerate(model.model.models):
        m.covar_module.raw_outputscale.data.fill_(float(oracle.gp["raw_outputscale"][c]))
        m.mean_module.constant.data.fill_(float(oracle.gp["constant"][c]))
        for nm in ("raw_variance", "raw_lengthscale", "raw_offset"):
            if nm in oracle.gp and hasattr(m.covar_module.base_kernel, nm):
                getattr(m.covar_module.base_kernel, nm).data.fill_(float(oracle.gp[nm][c]))


def model_cpu_sync(model, oracle, dev):
    """Re-synchronise the device model from the oracle's current parameters / buffers."""
    load_oracle_params(model, oracle)
    model._ensure_packed()


def rel_err(a, b):
    a, b = a.detach().cpu().double().reshape(-1), b.detach().cpu().double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_train_step(model_factory, dev, image_size=32, n_way=2, n_support=1, n_query=2, E=2, tol=1e-4,
                     steps=2, kernel="bncossim", lib=None, loose=None, report=None, same_branch=False,
                     gp_override=None):
    """Runs `steps` packed meta-train steps on the device path and on the oracle from identical weights
    and inputs; checks loss, every gradient, the post-Adam parameters, BN running statistics and the
    monitoring predictions (argmax bit-exact).  Tolerance: 1e-4 relative (north_star).

    Gradients, two modes.  Default (small shapes): device vs the free-running fp64 oracle, bar
    max(tol, k x the fp32 oracle's own distance to fp64).  ``same_branch=True`` (BASELINE shapes, where gate flips
    dominate -- tests/branch_checks.py): device vs a float64 replay of the step on the device's own ReLU / max-pool
    branch, bar `tol` flat; the free-running distances are still reported (keys ``free.*``, not asserted)."""
    from deep_kernel_transfer_b200.methods.DKT import DKT
    import branch_checks as bc
    torch.manual_seed(0)
    oracle = oep.OracleDKT("Conv4", kernel, n_way=n_way, n_support=n_support, seed=0,
                           feat_dim=64 * (image_size // 16) ** 2)
    # non-trivial GP hyper-parameters / affine parameters so that every gradient path is exercised
    oracle.gp["raw_outputscale"] = torch.linspace(-0.3, 0.6, n_way)
    oracle.gp["constant"] = torch.linspace(0.1, -0.2, n_way)
    for nm in ("raw_lengthscale", "raw_offset"):
        if nm in oracle.gp:
            oracle.gp[nm] = torch.linspace(1.5, 2.5, n_way) if nm == "raw_lengthscale" else torch.linspace(0.2, 0.6, n_way)
    if kernel == "linear":
        oracle.gp["raw_variance"] = torch.linspace(-0.2, 0.4, n_way)
    for k, v in (gp_override or {}).items():
        oracle.gp[k] = v.clone()
    g = torch.Generator().manual_seed(5)
    for k in oracle.bb:
        if k.endswith("BN.weight") or k.endswith("bn_out.weight"):
            oracle.bb[k] = 1.0 + 0.2 * torch.randn(oracle.bb[k].shape, generator=g)
        if k.endswith("BN.bias") or k.endswith("bn_out.bias"):
            oracle.bb[k] = 0.1 * torch.randn(oracle.bb[k].shape, generator=g)
    oracle64 = oep.OracleDKT("Conv4", kernel, n_way=n_way, n_support=n_support, seed=0, dtype=torch.float64,
                             feat_dim=64 * (image_size // 16) ** 2)
    model = DKT(model_factory, n_way, n_support, kernel=kernel, episodes_per_step=E, lib=lib)
    load_oracle_params(model, oracle)
    model = model.to(dev)
    model.train()
    worst = {}
    floor = {}
    info_only = set()
    branch_stats = {}

    def upd(key, val, val32=0.0):
        """val: device vs fp64 oracle; val32: fp32 oracle vs fp64 oracle (the reference's own rounding envelope)."""
        worst[key] = max(worst.get(key, 0.0), val)
        floor[key] = max(floor.get(key, 0.0), val32)

    for step in range(steps):
        xs = torch.stack([oep.synthetic_episode(100 * step + e, n_way, n_support, n_query, image_size)
                          for e in range(E)])
        # Parity is defined per step from IDENTICAL weights: Adam turns rounding-level
        # differences of near-zero gradients into +-lr moves, so trajectories are only comparable step-wise.
        oracle.optimizer = None
        if step > 0:
            for k in list(oracle.bb):
                oracle.bb[k] = oracle.bb[k].detach().clone()
            for k in list(oracle.gp):
                oracle.gp[k] = oracle.gp[k].detach().clone()
            model_cpu_sync(model, oracle, dev)
        snap = {k: v.detach().clone() for k, v in oracle.bb.items()}
        snap_gp = {k: v.detach().clone() for k, v in oracle.gp.items()}
        w_before = [b.C.weight.detach().cpu().clone() for b in model.feature.blocks()]
        oracle64.optimizer = None
        oracle64.bb = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in oracle.bb.items()}
        oracle64.gp = {k: v.detach().double() for k, v in oracle.gp.items()}
        r64 = oracle64.train_step(xs.double(), monitor=False)
        ref = oracle.train_step(xs)
        model._ensure_packed()
        model._new_adam()
        mon_flag, model.monitor = model.monitor, False      # the monitoring pass re-uses the conv workspaces
        out = model.train_step(xs.to(dev))
        model.monitor = mon_flag
        assert int(out["info"].cpu().abs().sum()) == 0
        upd("loss", rel_err(out["loss"], r64["loss"]), rel_err(ref["loss"], r64["loss"]))
        got = bc.device_grads(model, kernel)
        if same_branch:
            gates = bc.collect_device_gates(model, snap, n_way * (n_support + n_query))
            loss_b, truth, _ = bc.replay_step64(xs, snap, snap_gp, gates, kernel, n_way, n_support, branch_stats)
            upd("loss.branch", rel_err(out["loss"], loss_b))
        else:
            truth = r64["grads"]
        names = {"raw_outputscale": "g.outputscale", "constant": "g.constant", "trunk.bn_out.weight": "g.bn_out.w",
                 "trunk.bn_out.bias": "g.bn_out.b"}
        for key, gdev in got.items():
            if key.endswith(".C.bias"):
                # the conv bias cancels inside BatchNorm: its gradient is rounding noise around zero;
                # only require it to be negligible w.r.t. the weight gradient scale
                scale = float(ref["grads"][key.replace(".C.bias", ".C.weight")].abs().max())
                assert float(gdev.abs().max()) <= 1e-3 * scale + 1e-6, key
                continue
            rk = names.get(key, "g." + key)
            f32 = rel_err(ref["grads"][key], r64["grads"][key])
            if same_branch:
                upd(rk, rel_err(gdev, truth[key]))                       # flat `tol` on the device's own branch
                upd("free." + rk, rel_err(gdev, r64["grads"][key]), f32)
                info_only.add("free." + rk)
            else:
                upd(rk, rel_err(gdev, truth[key]), f32)
        for i, b in enumerate(model.feature.blocks()):
            # first Adam step moves every weight by lr*sign(g) (|g| >> eps): compare where the sign is numerically
            # determined (and |g| >> Adam's eps = 1e-8: below that the update is lr*g/(|g|+eps), which
            # amplifies rounding-level differences of g -- small-gradient kernels such as rbf reach that regime)
            gref = truth["trunk.%d.C.weight" % i].float() if same_branch else ref["grads"]["trunk.%d.C.weight" % i]
            mask = (gref.abs() > 1e-3 * gref.abs().max()) & (gref.abs() > 1e-5)
            d_all = b.C.weight.detach().cpu() - w_before[i]
            if bool(mask.any()):
                if same_branch:
                    d_ref = (-1e-3 * gref / (gref.abs() + 1e-8))[mask]
                else:
                    d_ref = (oracle.bb["trunk.%d.C.weight" % i].detach() - w_before[i])[mask]
                upd("adam.w%d" % i, float((d_all[mask] - d_ref).abs().max() / 1e-3))
            # wiring of the fused Adam for EVERY element: first step from the device's own gradient
            gd = b.C.weight.grad.detach().cpu().double()
            expect = -1e-3 * gd / (gd.abs() + 1e-8)
            upd("adam.self%d" % i, float((d_all.double() - expect).abs().max() / 1e-3))
            upd("rv%d" % i, rel_err(b.BN.running_var, oracle.bb["trunk.%d.BN.running_var" % i]))
        # monitoring (steps 5-6): after Adam the two sides differ by +-lr sign flips of noise-level gradients
        # (conv biases), so the strict check re-synchronises the post-update weights from the oracle and
        # re-runs the device monitoring, which still holds the pre-update train-mode features
        C, SQ = n_way, n_support + n_query
        model_cpu_sync(model, oracle, dev)
        mon_out = model.monitor_step(xs.to(dev))
        mean = mon_out["mean"].cpu().view(E, C, C, SQ)
        mean_s = mean[:, :, :, :n_support].reshape(E, C, C * n_support)
        mean_q = mean[:, :, :, n_support:].reshape(E, C, C * n_query)
        upd("mon.mean_s", rel_err(mean_s, ref["mean_support"]))
        upd("mon.mean_q", rel_err(mean_q, ref["mean_query"]))
        np.testing.assert_allclose(mon_out["acc_support"].cpu().numpy(), ref["acc_support"], atol=1e-4)
        np.testing.assert_allclose(mon_out["acc_query"].cpu().numpy(), ref["acc_query"], atol=1e-4)
        pred = mon_out["pred"].cpu().view(E, C, SQ)
        ref_ps = torch.sigmoid(ref["mean_support"]).numpy().argmax(axis=1).reshape(E, C, n_support)
        ref_pq = torch.sigmoid(ref["mean_query"]).numpy().argmax(axis=1).reshape(E, C, n_query)
        assert np.array_equal(pred[:, :, :n_support].numpy(), ref_ps), "support arg-max not bit-exact"
        assert np.array_equal(pred[:, :, n_support:].numpy(), ref_pq), "query arg-max not bit-exact"
    # bar: 1e-4 relative (north_star), or the fp32 reference's own distance to the fp64 truth where the tiny
    # test batches make BatchNorm's backward ill-conditioned (the device path must be as good as torch fp32)
    # The tcgen05 path multiplies in 3xTF32 (operands carried to ~2^-22 instead of fp32's 2^-24), so where the
    # envelope applies it is allowed 8x the fp32 reference's own distance to the fp64 truth instead of 3x.
    fac = 8.0 if torch.device(dev).type == "cuda" else 3.0
    # `loose`: {key prefix: tolerance} for quantities that sit behind ReLU / max-pool gates of a large batch, where a
    # single gate within fp32 rounding of its threshold flips between two correct evaluations
    def bar(k):
        if k in info_only:
            return float("inf")
        for pre, t in (loose or {}).items():
            if k.startswith(pre):
                return t
        return max(tol, fac * floor.get(k, 0.0))
    if report is not None:       # per key: (device vs fp64 truth, fp32 oracle vs fp64 oracle, bar applied)
        report.update({k: (v, floor.get(k, 0.0), bar(k)) for k, v in worst.items()})
        report["_branch"] = branch_stats
    bad = {k: (v, floor.get(k, 0.0)) for k, v in worst.items() if v > bar(k)}
    assert not bad, "parity above max(%g, %gx fp32-reference envelope): %s" % (tol, fac, bad)
    return model, oracle, worst


def check_correct(model, oracle, dev, image_size=32, n_way=2, n_support=1, n_query=3, tol=1e-4, extras=True,
                  episodes=2):
    model_cpu_sync(model, oracle, dev)      # identical weights / running statistics on both sides
    model.eval()
    for ep in range(episodes):
        x = oep.synthetic_episode(900 + ep, n_way, n_support, n_query, image_size)
        ref_logits = oracle.get_logits(x)
        got = model.get_logits(x)
        assert rel_err(got, ref_logits) <= tol
        ref_c = oracle.correct(x)
        got_c = model.correct(x)
        assert got_c[:2] == ref_c[:2], (got_c, ref_c)
        assert np.array_equal(got.cpu().numpy().argmax(1), ref_logits.numpy().argmax(1))
    # E-packed test path: identical hit counts, one kernel chain; and test_loop (which packs) against the oracle's counts
    eps = [oep.synthetic_episode(900 + ep, n_way, n_support, n_query, image_size) for ep in range(episodes)]
    ref_hits = [oracle.correct(x)[0] for x in eps]
    hits = model.correct_packed(torch.stack(eps))
    assert [float(h) for h in hits.tolist()] == ref_hits, (hits, ref_hits)
    model.test_episodes_per_call = 3
    acc = model.test_loop([(x, None) for x in eps])
    assert abs(acc - float(np.mean([h / (n_way * n_query) * 100 for h in ref_hits]))) < 1e-9
    if not extras:
        return
    # Laplace branch: CUDA embeddings handed to scikit-learn, as the reference does (DKT.py:207-224)
    x = oep.synthetic_episode(940, n_way, n_support, n_query, image_size)
    assert model.correct(x, laplace=True) == oracle.correct_laplace(x)
    # test-time adaptation (DKT.correct with N > 0): 3 Adam steps on the GP hyper-parameters, which persist afterwards
    x = oep.synthetic_episode(950, n_way, n_support, n_query, image_size)
    ref_c = oracle.correct(x, N=3)
    got_c = model.correct(x, N=3)
    assert got_c[:2] == ref_c[:2], (got_c, ref_c)
    assert abs(got_c[2] - ref_c[2]) <= 2e-4 * max(1.0, abs(ref_c[2])), (got_c, ref_c)
    for nm in ("raw_outputscale", "constant"):
        mine = torch.stack([(m.covar_module.raw_outputscale if nm == "raw_outputscale" else m.mean_module.constant).detach().view(())
                            for m in model.model.models]).cpu()
        assert float((mine - oracle.gp[nm].detach().view(-1)).abs().max()) <= 2e-5, nm      # 3 steps of +-1e-3 each
    assert rel_err(model.get_logits(x), oracle.get_logits(x)) <= 10 * tol


def check_regression(dev, lib=None, image=36, n=7, n_support=4, tol=2e-4, kernel="rbf"):
    """DKT regression (Conv3 + RBF GP with learned noise): one train step (loss, every gradient) and one test episode
    (predictive mean, confidence region, MSE) against the oracle (fp64 arbiter for the gradients)."""
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT_regression import DKT as DKTR
    torch.manual_seed(0)
    # Conv3 at a reduced resolution for the CPU-emulation run: feature dim = 36 * side^2
    side = image
    for _ in range(3):
        side = (side - 5) // 2 + 1
    feat = 36 * side * side
    o32 = oep.OracleDKTRegression(kernel, seed=0, feat_dim=feat)
    gq = torch.Generator().manual_seed(9)
    if kernel == "rbf":
        o32.gp["raw_lengthscale"] = torch.tensor([1.2])
        o32.gp["raw_outputscale"] = torch.tensor([0.4])
    else:       # small frequencies / scales so that the product over all feature dimensions does not underflow
        o32.gp["raw_mixture_weights"] = torch.randn(4, generator=gq) * 0.3
        o32.gp["raw_mixture_means"] = torch.randn(4, 1, feat, generator=gq) * 0.3 - 4.0
        o32.gp["raw_mixture_scales"] = torch.randn(4, 1, feat, generator=gq) * 0.3 - 4.0
    o32.gp["constant"] = torch.tensor([0.1])
    o32.gp["raw_noise"] = torch.tensor([-1.0])
    o64 = oep.OracleDKTRegression(kernel, seed=0, dtype=torch.float64, feat_dim=feat)
    o64.gp = {k: v.double() for k, v in o32.gp.items()}
    model = DKTR(backbone.Conv3(), kernel=kernel, lib=lib, feat_dim=feat)
    sd = {k: v.clone() for k, v in o32.bb.items()}
    model.feature_extractor.load_state_dict(sd)

    def sync_gp(o):
        cm = model.model.covar_module
        if kernel == "rbf":
            cm.base_kernel.raw_lengthscale.data.fill_(float(o.gp["raw_lengthscale"].detach()))
            cm.raw_outputscale.data.fill_(float(o.gp["raw_outputscale"].detach()))
        else:
            for nm in ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"):
                getattr(cm, nm).data.copy_(o.gp[nm].detach().float().to(getattr(cm, nm).device))
        model.model.mean_module.constant.data.fill_(float(o.gp["constant"].detach()))
        model.likelihood.noise_covar.raw_noise.data.fill_(float(o.gp["raw_noise"].detach()))
    sync_gp(o32)
    model = model.to(dev)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, 3, image, image, generator=g)
    y = torch.linspace(-1, 1, n)
    r32 = o32.train_step(x, y)
    r64 = o64.train_step(x.double(), y.double())
    loss = model.train_step(x.to(dev), y.to(dev))
    assert int(model._last_info.cpu().abs().sum()) == 0
    assert rel_err(loss, r64["loss"]) <= 1e-4
    got = {"layer%d.weight" % (i + 1): l.weight.grad for i, l in enumerate(model.feature_extractor.layers())}
    got.update({"layer%d.bias" % (i + 1): l.bias.grad for i, l in enumerate(model.feature_extractor.layers())})
    got["constant"] = model.model.mean_module.constant.grad.view(1)
    got["raw_noise"] = model.likelihood.noise_covar.raw_noise.grad.view(1)
    if kernel == "rbf":
        got["raw_outputscale"] = model.model.covar_module.raw_outputscale.grad.view(1)
        got["raw_lengthscale"] = model.model.covar_module.base_kernel.raw_lengthscale.grad.view(1)
    else:
        for nm in ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"):
            got[nm] = getattr(model.model.covar_module, nm).grad
    bad = {}
    for k, v in got.items():
        e = rel_err(v, r64["grads"][k])
        floor = rel_err(r32["grads"][k], r64["grads"][k])
        if e > max(tol, 3.0 * floor):
            bad[k] = (e, floor)
    assert not bad, bad
    # test episode: sync the (pre-step) parameters again, fit on a support subset, predict all
    o = oep.OracleDKTRegression(kernel, seed=0, feat_dim=feat)
    o.gp = {k: v.detach().clone() for k, v in o32.gp.items()}
    model.feature_extractor.load_state_dict({k: v.detach().clone() for k, v in o.bb.items()})
    sync_gp(o)
    idx = [0, 2, 3, 5][:n_support]
    mse_ref, mean_ref, lo_ref, hi_ref = o.test_episode(x[idx], y[idx], x, y)
    pred = model.predict(x[idx].to(dev), y[idx].to(dev), x.to(dev))
    lo, hi = pred.confidence_region()
    assert rel_err(pred.mean, mean_ref) <= 1e-4
    assert rel_err(lo, lo_ref) <= 1e-4 and rel_err(hi, hi_ref) <= 1e-4
    assert rel_err(((pred.mean.cpu() - y) ** 2).mean(), mse_ref) <= 1e-4
    return model


def check_train_step_arch(arch, model_factory, dev, image_size, n_way=2, n_support=1, n_query=2, E=2, kernel="rbf",
                          lib=None, tol=1e-4, env_factor=4.0, grad_check=True, report=None, lengthscale=(1.5, 2.5),
                          same_branch=False, grad_tol=None):
    """One packed meta-train step of an arbitrary backbone (ResNet*) against the oracle: loss, every backbone / GP
    gradient (fp64 arbiter + fp32 envelope), monitoring arg-max after re-synchronising the post-step weights."""
    from deep_kernel_transfer_b200.methods.DKT import DKT
    torch.manual_seed(0)
    o32 = oep.OracleDKT(arch, kernel, n_way=n_way, n_support=n_support, seed=0)
    o32.gp["raw_outputscale"] = torch.linspace(-0.3, 0.6, n_way)
    o32.gp["constant"] = torch.linspace(0.1, -0.2, n_way)
    if "raw_lengthscale" in o32.gp:
        o32.gp["raw_lengthscale"] = torch.linspace(lengthscale[0], lengthscale[1], n_way)
    o64 = oep.OracleDKT(arch, kernel, n_way=n_way, n_support=n_support, seed=0, dtype=torch.float64)
    o64.bb = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in o32.bb.items()}
    o64.gp = {k: v.detach().double() for k, v in o32.gp.items()}
    model = DKT(model_factory, n_way, n_support, kernel=kernel, episodes_per_step=E, lib=lib)
    sd = model.feature.state_dict()
    model.feature.load_state_dict({k: o32.bb[k].detach().clone() for k in sd})
    for c, m in enumerate(model.model.models):
        m.covar_module.raw_outputscale.data.fill_(float(o32.gp["raw_outputscale"][c]))
        m.mean_module.constant.data.fill_(float(o32.gp["constant"][c]))
        if "raw_lengthscale" in o32.gp:
            m.covar_module.base_kernel.raw_lengthscale.data.fill_(float(o32.gp["raw_lengthscale"][c]))
    model = model.to(dev)
    model.train()
    xs = torch.stack([oep.synthetic_episode(e, n_way, n_support, n_query, image_size) for e in range(E)])
    gp0 = {k: v.detach().double().clone() for k, v in o32.gp.items()}      # pre-step hyper-parameters (Adam moves them)
    r64 = o64.train_step(xs.double(), monitor=False)
    ref = o32.train_step(xs)
    model._ensure_packed()
    model._new_adam()
    snap = {k: v.detach().clone() for k, v in model.feature.state_dict().items()}
    eng = model.feature.engine(image_size, dev, lib)
    eng.keep_tape = bool(same_branch)
    model.monitor = False
    out = model.train_step(xs.to(dev))
    model.monitor = True
    assert int(out["info"].cpu().abs().sum()) == 0
    assert rel_err(out["loss"], r64["loss"]) <= tol, rel_err(out["loss"], r64["loss"])
    if report is not None:
        report["loss"] = (rel_err(out["loss"], r64["loss"]), rel_err(ref["loss"], r64["loss"]), tol)
    if same_branch:
        # every gradient against a float64 replay of the whole step (op tape -> head -> -mll) on the DEVICE's branch
        # (stored ReLU outputs / max-pool taps), flat `tol`; the replay runs in float64 on the same device
        from oracle import gp as ogp
        import torch.nn.functional as F
        name_of = {id(p_): n for n, p_ in model.feature.named_parameters()}
        leaves = {}

        def leaf(t):
            n = name_of[id(t)]
            if n not in leaves:
                leaves[n] = snap[n].detach().double().clone().requires_grad_(True)
            return leaves[n]
        stats = {"relu_flips": 0, "pool_flips": 0, "gates": 0}
        feats64 = resnet_replay64(eng.last_tape, True, stats, leaf)
        feats64.retain_grad()
        gp64 = {k: v.clone().requires_grad_(k in ogp.trainable_gp_names(kernel)) for k, v in gp0.items()}
        N = n_way * (n_support + n_query)
        tg = oep.make_targets(n_way, n_support + n_query, torch.float64)
        tot = 0.0
        for e in range(E):
            f = feats64[e * N:(e + 1) * N].cpu()            # the GP head of the replay runs on the host (oracle/gp.py)
            if kernel in oep.NORMALIZED_KERNELS:
                f = F.normalize(f, p=2, dim=1)
            tot = tot + ogp.mll_loss(kernel, f, tg, gp64) / E
        tot.backward()
        # the same replay in float32 (torch / cuDNN on the same gates): what fp32 arithmetic itself loses through the depth
        # of the network on a FIXED branch -- the envelope for the device's fp32-class arithmetic (6x: 3xTF32 operands carried
        # to 2^-21 instead of 2^-24, tensor-core accumulation that does not round to nearest)
        leaves32 = {}

        def leaf32(t):
            n = name_of[id(t)]
            if n not in leaves32:
                leaves32[n] = snap[n].detach().float().clone().requires_grad_(True)
            return leaves32[n]
        stats32 = {"relu_flips": 0, "pool_flips": 0, "gates": 0}
        tf32_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # true fp32, not TF32
        try:
            f32 = resnet_replay64(eng.last_tape, True, stats32, leaf32, dtype=torch.float32)
            f32.backward(feats64.grad.float())
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_flags
        floor_b = {n: rel_err(leaves32[n].grad, leaves[n].grad) for n in leaves if leaves[n].grad is not None
                   and n in leaves32 and leaves32[n].grad is not None}
        badb = {}
        for n, p_ in model.feature.named_parameters():
            if n not in leaves or leaves[n].grad is None:
                continue
            wname = n.replace(".bias", ".weight")
            if n.endswith(".bias") and p_.dim() == 1 and wname in leaves and \
                    leaves[n].grad.abs().max() < 1e-6 * leaves[wname].grad.abs().max():
                continue      # a bias that BatchNorm cancels
            e_ = rel_err(p_.grad, leaves[n].grad)
            bar_ = max(grad_tol or tol, 6.0 * floor_b.get(n, 0.0))      # 6 x what float32 itself loses on this branch
            if report is not None:
                report["g." + n] = (e_, floor_b.get(n, 0.0), bar_)
                report["free.g." + n] = (rel_err(p_.grad, r64["grads"][n]), rel_err(ref["grads"][n], r64["grads"][n]), float("inf"))
            if e_ > bar_:
                badb[n] = (e_, floor_b.get(n, 0.0))
        # hyper-parameter gradients: no gate in between, but sums over alpha = K~^-1 (y - m) amplify any difference of the
        # FEATURES by ~cond(K~) (N s / sigma^2 ~ 1e3 with an RBF kernel on un-normalised features).  Chain rule, link by
        # link: (i) device features vs the float64 replay on the device branch: 1e-4; (ii) the GP's hyper-gradients vs
        # the float64 GP evaluated AT the device's features: 1e-4.  The end-to-end distance is reported (free.*).
        feats_dev = eng.last_tape[-1][2].detach().double().cpu()
        e_feat = rel_err(feats_dev, feats64)
        if report is not None:
            report["features.branch"] = (e_feat, 0.0, tol)
        if e_feat > tol:
            badb["features"] = e_feat
        gpd = {k: v.clone().requires_grad_(k in ogp.trainable_gp_names(kernel)) for k, v in gp0.items()}
        totd = 0.0
        for e in range(E):
            f = feats_dev[e * N:(e + 1) * N]
            if kernel in oep.NORMALIZED_KERNELS:
                f = F.normalize(f, p=2, dim=1)
            totd = totd + ogp.mll_loss(kernel, f, tg, gpd) / E
        totd.backward()
        for nm, getter in (("raw_outputscale", lambda m: m.covar_module.raw_outputscale),
                           ("constant", lambda m: m.mean_module.constant),
                           ("raw_lengthscale", lambda m: m.covar_module.base_kernel.raw_lengthscale)):
            if nm in ogp.trainable_gp_names(kernel):
                gdev = torch.stack([getter(m).grad.view(()) for m in model.model.models])
                e_ = rel_err(gdev, gpd[nm].grad)
                if report is not None:
                    report["g." + nm] = (e_, 0.0, tol)
                    report["free.g." + nm] = (rel_err(gdev, gp64[nm].grad), rel_err(ref["grads"][nm], r64["grads"][nm]),
                                              float("inf"))
                if e_ > tol:
                    badb[nm] = e_
        if report is not None:
            report["_branch"] = stats
        assert not badb, ("same-branch gradients", badb)
        eng.last_tape = None
        eng.keep_tape = False
        grad_check = False       # the free-running comparison below is reported, not asserted
    bad = {}
    for name, p_ in model.feature.named_parameters():
        if name.endswith(".bias") and name.replace(".bias", ".weight") in ref["grads"] and p_.dim() == 1 and \
                ref["grads"][name].abs().max() < 1e-6 * ref["grads"][name.replace(".bias", ".weight")].abs().max():
            continue      # a bias that BatchNorm cancels: rounding noise on both sides
        e = rel_err(p_.grad, r64["grads"][name])
        floor = rel_err(ref["grads"][name], r64["grads"][name])
        if report is not None:
            report["g." + name] = (e, floor, max(tol, env_factor * floor))
        if grad_check and e > max(tol, env_factor * floor):
            bad[name] = (e, floor)
    gos = torch.stack([m.covar_module.raw_outputscale.grad for m in model.model.models])
    e = rel_err(gos, r64["grads"]["raw_outputscale"])
    if e > max(tol, env_factor * rel_err(ref["grads"]["raw_outputscale"], r64["grads"]["raw_outputscale"])):
        bad["raw_outputscale"] = e
    assert not bad, bad
    # monitoring with the oracle's post-step weights
    model.feature.load_state_dict({k: o32.bb[k].detach().clone() for k in sd})
    for c, m in enumerate(model.model.models):
        m.covar_module.raw_outputscale.data.fill_(float(o32.gp["raw_outputscale"][c].detach()))
        m.mean_module.constant.data.fill_(float(o32.gp["constant"][c].detach()))
        if "raw_lengthscale" in o32.gp:
            m.covar_module.base_kernel.raw_lengthscale.data.fill_(float(o32.gp["raw_lengthscale"][c].detach()))
    model._ensure_packed()
    mon = model.monitor_step(xs.to(dev))
    np.testing.assert_allclose(mon["acc_support"].cpu().numpy(), ref["acc_support"], atol=1e-4)
    np.testing.assert_allclose(mon["acc_query"].cpu().numpy(), ref["acc_query"], atol=1e-4)
    # predictive means of the monitoring pass: an OUTPUT, held to `tol` without any envelope
    C, SQ = n_way, n_support + n_query
    mean = mon["mean"].cpu().view(E, C, C, SQ)
    mean_s = mean[:, :, :, :n_support].reshape(E, C, C * n_support)
    mean_q = mean[:, :, :, n_support:].reshape(E, C, C * n_query)
    e_mean = max(rel_err(mean_s, ref["mean_support"]), rel_err(mean_q, ref["mean_query"]))
    if report is not None:
        report["mon.mean"] = (e_mean, 0.0, tol)
    assert e_mean <= tol, ("monitoring predictive mean", e_mean)
    return model


def check_sines(dev, lib=None, steps=3, tol=2e-4):
    """sines/train_DKT.py protocol: per-step loss / gradients / Adam update and prediction against the oracle."""
    from deep_kernel_transfer_b200.sines import SinesDKT, Task_Distribution
    np.random.seed(0)
    o = oep.OracleSines()
    o64 = oep.OracleSines(dtype=torch.float64)
    model = SinesDKT(lib=lib)
    model.net.load_state_dict({k: v.detach().clone() for k, v in o.p.items()})
    model = model.to(dev)
    tasks = Task_Distribution()
    for step in range(steps):
        x, y = tasks.sample_task().sample_data(10, noise=0.1)
        o64.p = {k: v.detach().double() for k, v in o.p.items()}
        o64.gp = {k: v.detach().double() for k, v in o.gp.items()}
        o64.optimizer = None
        r64 = o64.train_step(x.double(), y.double())
        ref = o.train_step(x, y)
        loss = model.train_step(x.to(dev), y.to(dev))
        assert int(model._last_info.cpu().abs().sum()) == 0
        assert rel_err(loss, r64["loss"]) <= tol
        got = {"layer1.weight": model.net.layer1.weight.grad, "layer1.bias": model.net.layer1.bias.grad,
               "layer2.weight": model.net.layer2.weight.grad, "layer2.bias": model.net.layer2.bias.grad,
               "constant": model.mean_module.constant.grad, "raw_noise": model.likelihood.noise_covar.raw_noise.grad,
               "raw_mixture_weights": model.covar_module.raw_mixture_weights.grad,
               "raw_mixture_means": model.covar_module.raw_mixture_means.grad,
               "raw_mixture_scales": model.covar_module.raw_mixture_scales.grad}
        bad = {}
        for k, v in got.items():
            e = rel_err(v, r64["grads"][k])
            floor = rel_err(ref["grads"][k], r64["grads"][k])
            if e > max(tol, 3.0 * floor):
                bad[k] = (e, floor)
        assert not bad, (step, bad)
        # re-synchronise so that the next step starts from identical weights (Adam sign flips of noise-level gradients)
        model.net.load_state_dict({k: v.detach().clone() for k, v in o.p.items()})
        model.mean_module.constant.data.copy_(o.gp["constant"].detach())
        model.likelihood.noise_covar.raw_noise.data.copy_(o.gp["raw_noise"].detach())
        for nm in ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"):
            getattr(model.covar_module, nm).data.copy_(o.gp[nm].detach().to(dev))
    x_all, y_all = tasks.sample_task().sample_data(40, noise=0.1, sort=True)
    mean_ref, var_ref = o.predict(x_all[:5], y_all[:5], x_all[5:])
    mean, var = model.predict(x_all[:5].to(dev), y_all[:5].to(dev), x_all[5:].to(dev))
    assert rel_err(mean, mean_ref) <= tol and rel_err(var, var_ref) <= tol


def resnet_replay64(tape, device_gates, stats, leaf, dtype=torch.float64):
    """Float64 replay of a ResNetEngine op tape (NCHW, torch autograd).  ``leaf(param_tensor)`` -> the float64 leaf to use
    for a module parameter.  device_gates: take the ReLU masks (y > 0 of the stored post-ReLU activations) and the
    max-pool taps (stored indices) from the device, asserting every disagreement with the free-running choice sits on a
    near-zero pre-activation / near-tie; otherwise run free.  Returns the features [B, D] (float64, differentiable)."""
    import torch.nn.functional as F

    def nchw(t):
        return t.detach().to(dtype).permute(0, 3, 1, 2)

    vals = {}
    vals[tape[0][1].data_ptr()] = nchw(tape[0][1])
    for rec in tape:
        if rec[0] == "conv":
            _, xi, out, m, (_, _, _, _, _, st, pad, dil) = rec
            bias = leaf(m.bias) if m.bias is not None else None
            vals[out.data_ptr()] = F.conv2d(vals[xi.data_ptr()], leaf(m.weight), bias, st, pad, dil)
        elif rec[0] == "bn":
            _, xi, y, m, _, _, res, relu, ipe = rec
            v = vals[xi.data_ptr()]
            o = torch.cat([F.batch_norm(v[e:e + ipe], None, None, leaf(m.weight), leaf(m.bias), True, 0.0, 1e-5)
                           for e in range(0, v.shape[0], ipe)])
            if res is not None:
                o = o + vals[res.data_ptr()]
            if relu and device_gates:
                mask = nchw(y) > 0
                diff = mask != (o > 0)
                stats["gates"] += mask.numel()
                if bool(diff.any()):
                    stats["relu_flips"] += int(diff.sum())
                    assert float(o[diff].abs().max()) <= 1e-4 * float(o.abs().max()), "ReLU gate differs off zero"
                o = o * mask
            elif relu:
                o = o.relu()
            vals[y.data_ptr()] = o
        elif rec[0] == "maxpool":
            _, xi, y, idx = rec
            v = vals[xi.data_ptr()]
            if device_gates:
                vp = F.pad(v, (1, 1, 1, 1), value=float("-inf"))
                pt = vp.unfold(2, 3, 2).unfold(3, 3, 2)                      # [B,C,Ho,Wo,3,3]
                pt = pt.reshape(*pt.shape[:4], 9)
                tap = idx.permute(0, 3, 1, 2).long().unsqueeze(-1)
                o = pt.gather(-1, tap).squeeze(-1)
                best = pt.max(-1).values
                bad = o != best
                if bool(bad.any()):
                    stats["pool_flips"] += int(bad.sum())
                    assert float((best - o)[bad].max()) <= 1e-4 * float(best.abs().max()), "max-pool tap differs off a tie"
            else:
                o = F.max_pool2d(v, 3, 2, 1)
            vals[y.data_ptr()] = o
        elif rec[0] == "avgpool":
            v = vals[rec[1].data_ptr()]
            vals[rec[2].data_ptr()] = F.avg_pool2d(v, v.shape[-1]).flatten(1)
    return vals[tape[-1][2].data_ptr()]


def check_resnet_same_branch(arch, dev, image_size, lib=None, B=4, ipe=4, tol=1e-4, fwd_tol=3e-5):
    """ResNet forward/backward at a given resolution against exact (fp64) arithmetic on the SAME piecewise-linear branch.

    A deep ReLU / max-pool network evaluated in fp32 picks a different gate than exact arithmetic wherever a
    pre-activation lies within rounding of zero (~10 of the 1e7 activations of ResNet18 at 224x224); each such flip moves
    every upstream gradient by ~1e-3 (measured: cuDNN's own fp32 path differs from fp64 by 1e-1 there), so raw
    gradients are only comparable on a fixed branch.  The check therefore (1) requires the device features to match a
    free-running fp64 replay of the op tape to `fwd_tol`, and that replay to match the oracle backbone (pinned to the
    reference's backbone.py) to 1e-10; (2) replays the tape in fp64 with the device's ReLU masks / arg-max taps,
    asserting every disagreeing gate sits on a near-zero pre-activation, and requires every parameter gradient to match
    that replay to `tol`."""
    import torch.nn.functional as F
    from deep_kernel_transfer_b200 import backbone, _lib
    from deep_kernel_transfer_b200.resnet_engine import ResNetEngine
    from oracle import backbone as obb
    lib = lib or _lib.load()
    dev = torch.device(dev)
    torch.manual_seed(0)
    net = getattr(backbone, arch)().to(dev)
    g = torch.Generator().manual_seed(3)
    for _, p in net.named_parameters():
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn(p.shape, generator=g).to(dev))
        p.grad = torch.zeros_like(p)
    eng = ResNetEngine(lib, net, dev)
    x = torch.randn(B, 3, image_size, image_size, generator=g).to(dev)
    feats = eng.forward(x, ipe, True)
    tape = list(eng.tape)
    gf = torch.randn(feats.shape, generator=g).to(dev)
    eng.backward(gf)
    stats = {"relu_flips": 0, "pool_flips": 0, "gates": 0}
    params = {}

    def leaf(t):
        if id(t) not in params:
            params[id(t)] = t.detach().double().clone().requires_grad_(True)
        return params[id(t)]

    def replay(device_gates):
        return resnet_replay64(tape, device_gates, stats, leaf), params

    f_free, _ = replay(False)
    sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
    f_or = obb.forward(arch, sd, x.double().cpu(), training=True, update_running=False)
    assert rel_err(f_free, f_or) <= 1e-10, ("tape replay vs oracle backbone", rel_err(f_free, f_or))
    assert rel_err(feats, f_free) <= fwd_tol, ("features", rel_err(feats, f_free))
    f_dev, params = replay(True)
    f_dev.backward(gf.double())
    bad = {}
    named = dict(net.named_parameters())
    for name, p in named.items():
        if id(p) in params and params[id(p)].grad is not None:
            wname = name.replace(".bias", ".weight")
            if name.endswith(".bias") and p.dim() == 1 and wname in named and id(named[wname]) in params and \
                    params[id(p)].grad.abs().max() < 1e-6 * params[id(named[wname])].grad.abs().max():
                continue      # a conv bias that the following BatchNorm cancels: rounding noise on both sides
            e = rel_err(p.grad, params[id(p)].grad)
            if e > tol:
                bad[name] = e
    assert not bad, bad
    assert stats["relu_flips"] + stats["pool_flips"] <= max(50, stats["gates"] // 100000), stats
    return stats


# ----------------------------------------------------------------------------- against the reference's own run
# tests/golden/dkt_*.npz: outputs of the reference's unmodified methods/DKT.py / DKT_regression.py
# (tests/golden/make_golden_dkt.py).  The drop-in module is driven through the SAME public calls.
import os

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Recorder:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, name, value, iteration):
        self.scalars.setdefault(name, []).append(float(value))

    def add_histogram(self, *a, **k):
        pass


def _golden_episodes(count, seed, n_way=3, per_class=5, image=84):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(count):
        x = torch.randn(n_way, per_class, 3, image, image, generator=g)
        out.append(x + 0.5 * torch.randn(n_way, 1, 3, 1, 1, generator=g))
    return out


def check_reference_golden_classification(dev, kernel, lib=None, tol=1e-4, drift_tol=5e-3):
    """train_loop / get_logits / correct / test_loop through the reference's call sequence (train.py:49-56,
    test.py:160-161) against what the reference's own DKT.py produced.  Step 1 of train_loop and the test path from the
    initial weights start from identical parameters: 1e-4 relative, arg-max results exact.  Later steps sit behind Adam
    (sign-like first updates amplify rounding-level gradients): `drift_tol`."""
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    gold = np.load(os.path.join(_GOLD, "dkt_cls_%s.npz" % kernel))
    oracle = oep.OracleDKT("Conv4", kernel, n_way=3, n_support=2, seed=5)      # parameter source only
    oracle.gp["constant"] = torch.tensor([0.05 * (c + 1) for c in range(3)])
    oracle.gp["raw_outputscale"] = torch.tensor([0.1 * c - 0.1 for c in range(3)])
    if kernel == "rbf":
        oracle.gp["raw_lengthscale"] = torch.tensor([30.0 + 5.0 * c for c in range(3)])
    if kernel == "linear":
        oracle.gp["raw_variance"] = torch.tensor([-3.0 + 0.3 * c for c in range(3)])
    model = DKT(backbone.Conv4, 3, 2, kernel=kernel, lib=lib)
    load_oracle_params(model, oracle)
    model = model.to(dev)
    test_eps = _golden_episodes(2, seed=13, per_class=6)
    model.eval()
    logits = torch.stack([model.get_logits(x).detach().cpu() for x in test_eps])
    assert rel_err(logits, torch.from_numpy(gold["logits_init"])) <= tol
    for e, x in enumerate(test_eps):
        model.n_query = x.size(1) - model.n_support
        assert tuple(model.correct(x)) == tuple(gold["correct_init"][e])
    acc_mean, acc_std = model.test_loop([(x, None) for x in test_eps], return_std=True)
    acc = gold["correct_init"][:, 0] / gold["correct_init"][:, 1] * 100
    assert abs(acc_mean - acc.mean()) < 1e-9 and abs(acc_std - acc.std()) < 1e-9
    model.train()
    model.writer = _Recorder()
    model.train_loop(0, [(x, None) for x in _golden_episodes(3, seed=11)], None, print_freq=1)
    loss = np.array(model.writer.scalars["loss"])
    assert abs(loss[0] - gold["loss"][0]) <= tol * abs(gold["loss"][0]), (loss, gold["loss"])
    assert model.writer.scalars["GP_support_accuracy"][0] == gold["acc_support"][0]
    assert model.writer.scalars["GP_query_accuracy"][0] == gold["acc_query"][0]
    np.testing.assert_allclose(loss, gold["loss"], rtol=drift_tol)
    model.writer = _Recorder()
    model.train_loop(1, [(x, None) for x in _golden_episodes(1, seed=12)], None, print_freq=1)
    np.testing.assert_allclose(model.writer.scalars["loss"], gold["loss_second_call"], rtol=drift_tol)
    # after four Adam steps: hyper-parameters (lr 1e-4: +-1e-4 per step at most) and the trained test path
    for c, m in enumerate(model.model.models):
        assert abs(float(m.covar_module.raw_outputscale.detach()) - float(gold["after_gp_raw_outputscale"][c])) <= 2.5e-4
    model.eval()
    model.writer = None
    logits = torch.stack([model.get_logits(x).detach().cpu() for x in test_eps])
    assert rel_err(logits, torch.from_numpy(gold["logits"])) <= 10 * drift_tol
    return model


def check_reference_golden_regression(dev, kernel, lib=None, tol=1e-4, drift_tol=5e-3):
    """DKT_regression through the reference's calls (train_regression.py:33-41, test_regression.py:34)."""
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT_regression import DKT as DKTR
    gold = np.load(os.path.join(_GOLD, "dkt_reg_%s.npz" % kernel))
    g = torch.Generator().manual_seed(21)
    batch = torch.rand(3, 19, 3, 100, 100, generator=g)
    labels = torch.rand(3, 19, generator=g) * 2 - 1
    o = oep.OracleDKTRegression(kernel, seed=5)                                  # parameter source only
    model = DKTR(backbone.Conv3(), kernel=kernel, lib=lib, get_batch=lambda who: (batch.clone(), labels.clone()))
    model.feature_extractor.load_state_dict({k: v.clone() for k, v in o.bb.items()})
    model = model.to(dev)
    optimizer = torch.optim.Adam([{"params": model.model.parameters(), "lr": 0.001},
                                  {"params": model.feature_extractor.parameters(), "lr": 0.001}])
    losses = []
    step = model.train_step
    model.train_step = lambda a, b: (lambda v: (losses.append(float(v)), v)[1])(step(a, b))
    model.model.train(); model.feature_extractor.train(); model.likelihood.train()
    model.train_loop(1, optimizer)
    assert abs(losses[0] - gold["loss"][0]) <= tol * abs(gold["loss"][0]), (losses, gold["loss"])
    np.testing.assert_allclose(losses, gold["loss"], rtol=drift_tol)
    np.random.seed(3)
    mse = model.test_loop(5)
    assert abs(float(mse) - float(gold["test_mse"])) <= 10 * drift_tol * float(gold["test_mse"]), (float(mse), float(gold["test_mse"]))
    return model


# ----------------------------------------------------------------------------- sines: against the reference script's own run
# tests/golden/sines_reference.npz: observed while running the reference's unmodified sines/train_DKT.py::main()
# (tests/golden/make_golden_sines.py): data of the first training tasks, parameters before each of those steps, the
# losses it computed, the trained state after its 50 000 iterations, predictions on its first test tasks.
_SINES_GP = {"constant": "gp.mean_module.constant", "raw_noise": "gp.likelihood.noise_covar.raw_noise",
             "raw_mixture_weights": "gp.covar_module.raw_mixture_weights",
             "raw_mixture_means": "gp.covar_module.raw_mixture_means",
             "raw_mixture_scales": "gp.covar_module.raw_mixture_scales"}


def sines_gold():
    return np.load(os.path.join(_GOLD, "sines_reference.npz"))


def sines_state(gold, prefix):
    net = {k: torch.from_numpy(gold["%s.net.%s" % (prefix, k)].copy()) for k in
           ("layer1.weight", "layer1.bias", "layer2.weight", "layer2.bias")}
    gp = {k: torch.from_numpy(gold["%s.%s" % (prefix, v)].copy()) for k, v in _SINES_GP.items()}
    return net, gp


def sines_test_tasks(gold):
    """(x_support, y_support, x_query, mean, lower, upper) of the recorded test tasks; the support indices are recovered
    from the recorded support targets (exact float matches inside y_all)."""
    out = []
    j = 0
    while "test%d.mean" % j in gold.files:
        x_all, y_all = gold["test%d.x_all" % j], gold["test%d.y_all" % j]
        ys = gold["test%d.y_support" % j]
        idx = np.array([int(np.where(y_all == v)[0][0]) for v in ys])
        q = np.setdiff1d(np.arange(len(y_all)), idx)
        out.append((torch.from_numpy(x_all[idx]), torch.from_numpy(ys.copy()), torch.from_numpy(x_all[q]),
                    gold["test%d.mean" % j], gold["test%d.lower" % j], gold["test%d.upper" % j]))
        j += 1
    return out


def check_sines_oracle_against_reference_run(tol=2e-5):
    """oracle/episode.py::OracleSines vs the reference script: loss of each recorded step from the recorded
    parameters, the first Adam update, predictions of the trained model."""
    gold = sines_gold()
    n = len(gold["loss"])
    for i in range(n):
        o = oep.OracleSines()
        o.p, o.gp = sines_state(gold, "step%d" % i)
        r = o.train_step(torch.from_numpy(gold["train_x"][i]), torch.from_numpy(gold["train_y"][i]))
        assert abs(float(r["loss"]) - gold["loss"][i]) <= tol * abs(gold["loss"][i]), (i, float(r["loss"]), gold["loss"][i])
        if i == 0:      # a fresh Adam's first update: +-lr on every element whose gradient is not rounding noise
            nxt_net, nxt_gp = sines_state(gold, "step1")
            for k, v in o.p.items():
                big = r["grads"][k].abs() > 1e-6
                assert float((v.detach() - nxt_net[k]).abs()[big].max()) <= 2e-6, k
            for k, v in o.gp.items():
                if v.requires_grad:
                    big = r["grads"][k].abs() > 1e-6
                    assert float((v.detach() - nxt_gp[k].view_as(v)).abs()[big].max()) <= 2e-6, k
    o = oep.OracleSines()
    o.p, o.gp = sines_state(gold, "trained")
    mses = []
    for xs, ys, xq, mean, lower, upper in sines_test_tasks(gold):
        m, v = o.predict(xs, ys, xq)
        assert rel_err(m, torch.from_numpy(mean)) <= 1e-4
        sd2 = 2.0 * v.sqrt()
        assert rel_err(m - sd2, torch.from_numpy(lower)) <= 1e-4 and rel_err(m + sd2, torch.from_numpy(upper)) <= 1e-4
    return gold


def check_sines_against_reference_run(dev, lib=None, tol=1e-4):
    """The CUDA SinesDKT vs the reference script's own run: per-step loss from the recorded parameters, first Adam
    update, and the trained model's predictive mean / confidence region on the recorded test tasks."""
    from deep_kernel_transfer_b200.sines import SinesDKT
    gold = sines_gold()

    def load(model, prefix):
        net, gp = sines_state(gold, prefix)
        model.net.load_state_dict(net)
        with torch.no_grad():
            model.mean_module.constant.copy_(gp["constant"].view(1))
            model.likelihood.noise_covar.raw_noise.copy_(gp["raw_noise"].view(1))
            for nm in ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"):
                getattr(model.covar_module, nm).copy_(gp[nm].view_as(getattr(model.covar_module, nm)))
    for i in range(len(gold["loss"])):
        model = SinesDKT(lib=lib)
        load(model, "step%d" % i)
        model = model.to(dev)
        loss = model.train_step(torch.from_numpy(gold["train_x"][i]).to(dev), torch.from_numpy(gold["train_y"][i]).to(dev))
        assert int(model._last_info.cpu().abs().sum()) == 0
        assert abs(float(loss) - gold["loss"][i]) <= tol * abs(gold["loss"][i]), (i, float(loss), gold["loss"][i])
        if i == 0:
            nxt_net, _ = sines_state(gold, "step1")
            for k, v in model.net.state_dict().items():
                g = dict(model.net.named_parameters())[k].grad.cpu()
                big = g.abs() > 1e-5
                assert float((v.cpu() - nxt_net[k]).abs()[big].max()) <= 5e-6, k
    model = SinesDKT(lib=lib)
    load(model, "trained")
    model = model.to(dev)
    o32 = oep.OracleSines()
    o32.p, o32.gp = sines_state(gold, "trained")
    for xs, ys, xq, mean, lower, upper in sines_test_tasks(gold):
        m, v = model.predict(xs.to(dev), ys.to(dev), xq.to(dev))
        m, v = m.cpu(), v.cpu()
        # the goldens carry the GP arithmetic in float64; the trained spectral kernel evaluates cos(2 pi tau mu) over 40
        # dimensions, which fp32 itself only reproduces to ~1e-4: bar = max(tol, 3 x the fp32 oracle's own distance)
        m32, _ = o32.predict(xs, ys, xq)
        bar = max(tol, 3.0 * rel_err(m32, torch.from_numpy(mean)))
        assert rel_err(m, torch.from_numpy(mean)) <= bar, (rel_err(m, torch.from_numpy(mean)), bar)
        sd2 = 2.0 * v.sqrt()
        assert rel_err(m - sd2, torch.from_numpy(lower)) <= bar and rel_err(m + sd2, torch.from_numpy(upper)) <= bar
