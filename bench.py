#!/usr/bin/env python
"""Benchmark of the DKT meta-train hot path (BASELINE.json metric: episodes/sec, 5-way 5-shot Conv4
bncossim, synthetic 84x84x3 episodes).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                   # the reference's CPU path (oracle port), rank 0 only

One "step" = one packed meta-step: every rank runs E episodes (weak scaling, fixed E per GPU) through the
full reference loop body (DKT.train_loop steps 1-6: train-mode forward over N=105 images, C=5 exact-GP
marginal likelihoods, backward, Adam, and the eval-mode monitoring forward + predictive means + arg-max),
followed under torch.distributed by one all-reduce of the flat gradient buffer.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic MACs per image (measured from the reference's backbone.py): whole backbone / first conv
MACS = {"Conv4": (97164288, 12192768), "ResNet18": (1813561344, 118013952), "ResNet50": (4087136256, 118013952)}

# BASELINE.json configs -> what one "step" is.  cfg3 (configs[2]) is the configuration the metric is quoted on: default.
CONFIGS = {
    "cfg3": dict(arch="Conv4", kernel="bncossim", n_way=5, n_support=5, n_query=16, image=84, feat=1600, E=32, graph=False,
                 workload="5-way 5-shot Conv4 bncossim, synthetic 84x84x3, meta-train step incl. monitoring (DKT.train_loop body)"),
    "cfg2": dict(arch="Conv4", kernel="bncossim", n_way=5, n_support=1, n_query=16, image=84, feat=1600, E=1, graph=True,
                 workload="5-way 1-shot Conv4 bncossim, synthetic 84x84x3, meta-train step incl. monitoring, ONE episode per "
                          "Adam step (the reference's own granularity), step replayed from a CUDA graph"),
    "cfg4": dict(arch="ResNet18", kernel="rbf", n_way=5, n_support=5, n_query=16, image=224, feat=512, E=4, graph=False,
                 workload="5-way 5-shot ResNet18 RBF, synthetic 224x224x3, meta-train step incl. monitoring"),
    "cfg5-e2e": dict(arch="ResNet50", kernel="bncossim", n_way=20, n_support=5, n_query=16, image=224, feat=2048, E=1, graph=False,
                     workload="20-way 5-shot ResNet50 bncossim (the reference's default kernel), synthetic 224x224x3 (N = 420 "
                              "exact-GP systems), meta-train step incl. monitoring"),
}
CFG = CONFIGS["cfg3"]
N_WAY, N_SUPPORT, N_QUERY, IMAGE = 5, 5, 16, 84
WORKLOAD = CFG["workload"]
M_BB, M_FIRST = MACS["Conv4"]


def select_config(name):
    global CFG, N_WAY, N_SUPPORT, N_QUERY, IMAGE, WORKLOAD, M_BB, M_FIRST
    CFG = CONFIGS[name]
    N_WAY, N_SUPPORT, N_QUERY, IMAGE = CFG["n_way"], CFG["n_support"], CFG["n_query"], CFG["image"]
    WORKLOAD = CFG["workload"]
    M_BB, M_FIRST = MACS[CFG["arch"]]


def episode_flops(monitor=True):
    """Algorithmic work per training episode: N*(6*M_bb - 2*M_first) + monitoring 2*M_bb*N + GP."""
    n, d, c = N_WAY * (N_SUPPORT + N_QUERY), CFG["feat"], N_WAY
    f = n * (6 * M_BB - 2 * M_FIRST) + 4 * n * n * d + c * n ** 3
    if monitor:
        f += 2 * M_BB * n
    return f


def make_model(E, dev=None):
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    m = DKT(getattr(backbone, CFG["arch"]), N_WAY, N_SUPPORT, kernel=CFG["kernel"], episodes_per_step=E)
    if CFG["kernel"] == "rbf":      # un-normalised features: a lengthscale that keeps |x - x'|^2 / l^2 = O(1)
        import math
        for mm in m.model.models:
            mm.covar_module.base_kernel.raw_lengthscale.data.fill_(math.sqrt(CFG["feat"]))
    return m.to(dev) if dev is not None else m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(n_timed=3, n_warm=1):
    """The reference's CPU path (oracle port: reference backbone.py semantics + restated GPyTorch math) on the
    host cores: full train_loop body per episode, E=1 (the reference's own granularity).  ResNet configs run a BOUNDED
    sample: the backbone work of `frac` of the episode's images (forward + backward + monitoring forward scale linearly
    with the image count) plus the full-size GP on features of the right shape, added up and reported as such."""
    import torch
    from oracle import episode as oep
    from oracle import gp as ogp
    cores = os.cpu_count() or 1
    arch, kernel = CFG["arch"], CFG["kernel"]
    # pick the intra-op thread count that runs the backbone fastest on this host (large shared hosts are
    # slower with one thread per logical core); the chosen count is reported as `cores`
    n_probe = N_WAY * (N_SUPPORT + N_QUERY) if arch == "Conv4" else 8
    xprobe = oep.synthetic_episode(0, N_WAY, N_SUPPORT, N_QUERY, IMAGE).reshape(-1, 3, IMAGE, IMAGE)[:n_probe]
    probe = oep.OracleDKT(arch, kernel, n_way=N_WAY, n_support=N_SUPPORT, seed=0)
    best = (1e30, cores)
    for th in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
        torch.set_num_threads(th)
        with torch.no_grad():
            oep.features(arch, probe.bb, xprobe, kernel, training=False)
            t0 = time.perf_counter()
            oep.features(arch, probe.bb, xprobe, kernel, training=False)
            dt = time.perf_counter() - t0
        best = min(best, (dt, th))
    threads = best[1]
    torch.set_num_threads(threads)
    if arch == "Conv4":
        o = oep.OracleDKT(arch, kernel, n_way=N_WAY, n_support=N_SUPPORT, seed=0)
        ts = []
        for i in range(n_warm + n_timed):
            x = oep.synthetic_episode(i, N_WAY, N_SUPPORT, N_QUERY, IMAGE)
            t0 = time.perf_counter()
            o.train_step(x, monitor=True)
            dt = time.perf_counter() - t0
            if i >= n_warm:
                ts.append(dt)
        ts.sort()
        med = ts[len(ts) // 2]
        sample = "%d warm-up + %d timed single-episode meta-train steps (median), same synthetic episodes" % (n_warm, n_timed)
    else:
        # bounded sample: per-class shots reduced so that the CPU work stays ~10-30 s; backbone time scales with images
        n_full = N_WAY * (N_SUPPORT + N_QUERY)
        q_small = 1 if arch == "ResNet50" else 3
        o = oep.OracleDKT(arch, kernel, n_way=N_WAY, n_support=1, seed=0)
        o.gp["raw_lengthscale"] = torch.full((N_WAY,), float(CFG["feat"]) ** 0.5)
        x = oep.synthetic_episode(0, N_WAY, 1, q_small, IMAGE)
        n_small = N_WAY * (1 + q_small)
        t0 = time.perf_counter()
        o.train_step(x, monitor=True)
        t_small = time.perf_counter() - t0
        # the GP of the full-size episode on features of the right shape (the small episode's GP is negligible)
        z = torch.randn(n_full, CFG["feat"]).abs().requires_grad_(True)
        p = {k: v.detach().clone().requires_grad_(k in ogp.trainable_gp_names(kernel)) for k, v in o.gp.items()}
        tg = oep.make_targets(N_WAY, N_SUPPORT + N_QUERY)
        t0 = time.perf_counter()
        ogp.mll_loss(kernel, z, tg, p).backward()
        with torch.no_grad():
            ogp.predict(kernel, z.detach(), tg, z.detach(), {k: v.detach() for k, v in p.items()})
        t_gp = time.perf_counter() - t0
        med = t_small * n_full / n_small + t_gp
        sample = ("bounded: one %d-image episode of the same backbone / resolution (%.1f s) scaled by %d/%d images + the "
                  "full-size GP (N = %d, C = %d: %.1f s) on features of the right shape"
                  % (n_small, t_small, n_full, n_small, n_full, N_WAY, t_gp))
    return {"value": 1.0 / med, "unit": "episodes/s", "cores": threads, "host_logical_cores": cores, "kind": "port",
            "sample": sample, "s_per_episode": med}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb = cpu_baseline(n_timed=steps, n_warm=warm)
    line = {"metric": "episodes/sec (meta-train)", "value": cb["value"], "unit": "episodes/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * cb["s_per_episode"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "episodes_per_step": 1, "note":
                       "reference CPU path = oracle port (GPyTorch is not installable offline; backbone pinned to the "
                       "reference's backbone.py); each step is one episode, the reference's own granularity"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--episodes-per-gpu", type=int, default=None, help="episodes per GPU per step (default: the config's)")
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS) + ["cfg5-sweep", "test"],
                    help="BASELINE.json configs: cfg3 = configs[2] (the metric's configuration, default), cfg2 = configs[1], "
                         "cfg4 = configs[3], cfg5-sweep / cfg5-e2e = configs[4], test = the test path (DKT.correct / test_loop)")
    ap.add_argument("--no-monitor", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="train", choices=["train", "feeder"],
                    help="train = the BASELINE metric (default); feeder = the GPU episode feeder alone (SURVEY 8f-1)")
    args = ap.parse_args()
    if args.config in CONFIGS:
        select_config(args.config)
    if args.episodes_per_gpu is None:
        args.episodes_per_gpu = CFG["E"] if args.config in CONFIGS else 32
    if args.workload == "feeder":
        return run_feeder(args)
    if args.config == "cfg5-sweep":
        return run_gp_sweep(args)
    if args.config == "test":
        return run_test_path(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from deep_kernel_transfer_b200 import _lib, backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    from oracle import episode as oep      # synthetic-episode generator only (no arithmetic of the product path)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # feeder placement: run (and first-touch the pinned episode buffers) on the GPU's NUMA node
    from deep_kernel_transfer_b200.feeder import bind_to_gpu_numa
    affinity0 = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    E = args.episodes_per_gpu
    W = max(3, args.warmup)
    K = args.steps
    torch.manual_seed(0)
    model = make_model(E, dev)
    model.cuda_graph = bool(CFG["graph"]) and world == 1
    model.monitor = not args.no_monitor
    model.train()
    model._ensure_packed()
    model._new_adam()
    # synthetic episodes: a small pool of distinct seeded episodes per rank, tiled to E (content does not
    # change the work); the pinned host copy feeds the end-to-end leg
    pool = torch.stack([oep.synthetic_episode(1000 * rank + i, N_WAY, N_SUPPORT, N_QUERY, IMAGE) for i in range(4)])
    host = pool[torch.arange(E) % 4].contiguous().pin_memory()
    x_dev = host.to(dev)
    step_bytes_in = host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: inputs already in HBM
    for _ in range(W):
        model.train_step(x_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = lib.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(K):
        out = model.train_step(x_dev)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = lib.launches - l0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    info_bad = int((out["info"] != 0).sum().item())
    # ---------------- end-to-end leg: pinned host input -> H2D -> step -> D2H of the losses, every step
    from deep_kernel_transfer_b200.feeder import DevicePrefetcher
    for _ in range(2):
        model.train_step(host.to(dev, non_blocking=True))["loss"].cpu()
    barrier()
    # the same path DKT.train_loop uses: every step's input starts in pinned host memory, its H2D copy is issued
    # inside the timed region (overlapped with the previous step's kernels) and the step's results are read back.
    # Buffers / streams are set up before the clock starts; every copy (including step 0's) is inside it.
    feed = DevicePrefetcher((host for _ in range(K)), dev, timing=True, start=False)
    feed.preallocate(host.shape)
    n_res = (3 if model.monitor else 1) * E
    d2h_stream = torch.cuda.Stream(dev)
    host_res = [torch.empty(n_res, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending, marks, last_loss = None, [], None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    feed.start()
    for k, xb in enumerate(feed):
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        o = model.train_step(xb)
        m1.record()
        marks.append((m0, m1))
        feed.release(xb)
        cur = torch.cat([o["loss"], o["acc_support"], o["acc_query"]] if model.monitor else [o["loss"]])
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(d2h_stream):        # the read-back of step k must not queue behind step k+1's kernels
            d2h_stream.wait_event(ready)
            host_res[k & 1].copy_(cur, non_blocking=True)
            done = torch.cuda.Event()
            done.record(d2h_stream)
        cur.record_stream(d2h_stream)
        if pending is not None:                    # the host consumes step k-1's losses / accuracies while step k runs
            pending[0].synchronize()
            last_loss = float(pending[1][:E].mean())
        pending = (done, host_res[k & 1])
    pending[0].synchronize()
    last_loss = float(pending[1][:E].mean())
    d2h = n_res * 4
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    first_start_ms = e0.elapsed_time(marks[0][0])
    step_ms = sorted(a.elapsed_time(b) for a, b in marks)
    gap_ms = sorted(marks[i][1].elapsed_time(marks[i + 1][0]) for i in range(len(marks) - 1))
    h2d_ms = sorted(a.elapsed_time(b) for a, b in feed.copy_events)
    h2d_ms_med = h2d_ms[len(h2d_ms) // 2]
    # ---------------- feeder-fed leg: the same steps fed by the GPU episode feeder from a resident uint8 store
    # (host draws the episode composition + augmentation parameters, one transform kernel per step)
    ms_feed, feed_info = float("nan"), None
    if args.config == "cfg3":
        try:
            ms_feed, feed_info = feeder_fed_leg(model, lib, dev, E, K, rank)
        except Exception as ex:      # the headline legs above stand on their own
            feed_info = {"error": repr(ex)}
    barrier()
    t = torch.tensor([ms, ms_e2e, ms_feed], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_feed = float(t[0]), float(t[1]), float(t[2])
    # ---------------- roofline of the dominant kernel (64->64 3x3 conv at 42x42), timed live on its stream
    roof = None
    if rank == 0:
        roof = dominant_kernel_roofline(model, lib, dev, E) if CFG["arch"] == "Conv4" else resnet_roofline(model, lib, dev, E)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    eps = E * world * K / (ms / 1000.0)
    eps_e2e = E * world * K / (ms_e2e / 1000.0)
    line = {
        "metric": "episodes/sec (meta-train)", "value": eps, "unit": "episodes/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_way": N_WAY, "n_shot": N_SUPPORT, "n_query": N_QUERY, "image": IMAGE,
                   "episodes_per_gpu_per_step": E, "global_episodes_per_step": E * world, "monitor": model.monitor,
                   "parallelism": "episodes sharded dp%d, one all-reduce of the flat gradient per step" % world,
                   "cuda_graph": bool(model.cuda_graph), "bench_config": args.config,
                   "l2": "per-step working set (inputs %.0f MB + activations) exceeds the 126 MB L2; every step re-reads its "
                         "inputs from HBM" % (step_bytes_in / 1e6)},
        "gpu_launches": launches,
        "tflops_algorithmic": eps * episode_flops(monitor=model.monitor) / 1e12,
        "cholesky_failures": info_bad,
        "e2e": {"value": eps_e2e, "unit": "episodes/s", "h2d_bytes_per_step": step_bytes_in,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                "h2d_ms_per_step": h2d_ms_med, "h2d_gb_per_s": step_bytes_in / h2d_ms_med / 1e6,
                "step_kernels_ms_median": step_ms[len(step_ms) // 2], "step_kernels_ms_max": step_ms[-1],
                "inter_step_gap_ms_median": gap_ms[len(gap_ms) // 2] if gap_ms else 0.0,
                "inter_step_gap_ms_max": gap_ms[-1] if gap_ms else 0.0, "first_step_start_ms": first_start_ms,
                "note": "H2D of step k+1 overlaps step k on a copy stream; when the host link moves the step's input slower "
                        "than one step computes, the end-to-end rate is the link's"},
        "clocks": clocks, "roofline": roof,
    }
    line["e2e"]["feeder_cpus"] = len(numa_cpus) if numa_cpus else None
    if feed_info is not None and "error" not in feed_info:
        feed_info.update({"value": E * world * K / (ms_feed / 1000.0), "unit": "episodes/s", "ms_per_step": ms_feed / K})
    line["e2e_device_feeder"] = feed_info
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)          # the CPU baseline may use every host core
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def feeder_fed_leg(model, lib, dev, E, K, rank):
    """K meta-train steps whose episodes are assembled on the device by the episode feeder (RandomSizedCrop + jitter
    + flip + normalise of uint8 CUB-sized images resident in HBM).  Returns (ms over the K steps, info dict)."""
    import numpy as np
    import torch
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore, EpisodeFeeder
    rs = np.random.RandomState(rank)
    shapes = [(375, 500), (333, 500), (500, 375), (400, 500), (281, 500), (500, 500), (357, 500), (500, 332)]
    n_classes, per_class = 100, 30
    hw = [shapes[int(rs.randint(len(shapes)))] for _ in range(n_classes * per_class)]
    store = EpisodeStore.from_device_bytes(hw, np.repeat(np.arange(n_classes), per_class), dev, seed=rank)
    feed = EpisodeFeeder(store, IMAGE, N_WAY, N_SUPPORT, N_QUERY, n_episode=2 * E, aug=True, seed=rank, lib=lib)
    for x in feed.device_packs(E):                     # warm-up (allocations, kernel attribute)
        model.train_step(x)
    torch.cuda.synchronize()
    feed.n_episode = K * E
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d, loss = 0, None
    e0.record()
    for x in feed.device_packs(E):
        o = model.train_step(x)
        h2d += feed.last_params[0].numel() * 4 + feed.last_params[1].numel() * 4
        loss = o["loss"]
    float(loss.mean())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), {"h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": 4,
                                 "store_images": len(store), "store_gb": store.nbytes / 1e9,
                                 "note": "train_loop fed by EpisodeFeeder.device_packs: 44 B of parameters per image "
                                         "cross the host link instead of 84.7 KB of fp32 pixels"}


def feeder_cpu_baseline(images, params, factors, size):
    """The reference's own transform chain (PIL + torchvision, data/datamgr.py:37-46) on one host core over a bounded
    sample of the same images and augmentation parameters; falls back to the numpy oracle port if PIL is missing."""
    import numpy as np
    import torch
    t0 = time.perf_counter()
    try:
        from PIL import Image, ImageEnhance
        import torchvision.transforms.functional as TF
        from torchvision.transforms import InterpolationMode
        kind = "reference"
        for img, p, f in zip(images, params, factors):
            o = TF.resized_crop(Image.fromarray(img), int(p[1]), int(p[2]), int(p[3]), int(p[4]), [size, size],
                                InterpolationMode.BILINEAR)
            for enh, a in zip((ImageEnhance.Brightness, ImageEnhance.Contrast, ImageEnhance.Color), f):
                o = enh(o).enhance(float(a)).convert("RGB")
            if p[5]:
                o = TF.hflip(o)
            TF.normalize(TF.to_tensor(o), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    except ImportError:
        from oracle import transforms as ot
        kind = "port"
        for img, p, f in zip(images, params, factors):
            ot.transform_aug(img, (int(p[1]), int(p[2]), int(p[3]), int(p[4])), f, int(p[5]), size)
    dt = time.perf_counter() - t0
    return kind, dt


def run_feeder(args):
    """The GPU episode feeder alone: E packed 5-way (5+16) episodes per step, augmentation on, built from a
    device-resident uint8 store of CUB-sized images.  value = kernel only (parameters resident); e2e = through
    EpisodeFeeder.device_packs (host sampling of the episode composition and augmentation parameters, their H2D copy,
    the kernel, and a D2H read of one value per step)."""
    import numpy as np
    import torch
    from deep_kernel_transfer_b200 import _lib
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore, EpisodeFeeder
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    E, K, W = args.episodes_per_gpu, args.steps, max(3, args.warmup)
    rs = np.random.RandomState(0)
    shapes = [(375, 500), (333, 500), (500, 375), (400, 500), (281, 500), (500, 500), (357, 500), (500, 332)]
    n_classes, per_class = 100, 30
    hw = [shapes[int(rs.randint(len(shapes)))] for _ in range(n_classes * per_class)]
    labels = np.repeat(np.arange(n_classes), per_class)
    store = EpisodeStore.from_device_bytes(hw, labels, dev, seed=0)
    feed = EpisodeFeeder(store, IMAGE, N_WAY, N_SUPPORT, N_QUERY, n_episode=E * (K + W), aug=True, seed=0, lib=lib)
    n_img = E * N_WAY * (N_SUPPORT + N_QUERY)
    out = torch.empty(n_img, 3, IMAGE, IMAGE, device=dev)
    draws = [feed.draw(E) for _ in range(4)]
    for d in draws[:W]:
        feed.transform(d, out=out)
    feed.check()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    # kernel-only leg: CUDA events around each launch (the parameter upload of transform() precedes the first event)
    l0 = lib.launches
    t_wall0 = time.time()
    kernel_ms, alg_bytes = [], []
    for k in range(K):
        d = draws[k % len(draws)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        feed.transform(d, out=out, events=(e0, e1))
        torch.cuda.synchronize()
        kernel_ms.append(e0.elapsed_time(e1))
        p = d["params"]
        alg_bytes.append(float((p[:, 3].astype(np.int64) * p[:, 4] * 3).sum() + n_img * 3 * IMAGE * IMAGE * 4))
    launches = lib.launches - l0
    ms = float(sum(kernel_ms))
    # end-to-end leg through the public iterator
    feed.n_episode = E * K
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    h2d = 0
    for x in feed.device_packs(E):
        float(x[0, 0, 0, 0, 0, 0])
        h2d += feed.last_params[0].numel() * 4 + feed.last_params[1].numel() * 4
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6500.0)
    ach = sum(alg_bytes) / (ms / 1e3) / 1e9
    line = {"metric": "episodes/sec (episode feeder, aug)", "value": E * K / (ms / 1e3), "unit": "episodes/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "episode feeder: 5-way 5-shot + 16 queries, 84x84, RandomSizedCrop + ImageJitter + flip + "
                                   "normalise from a resident uint8 store of %d CUB-sized images (%.2f GB)"
                                   % (len(store), store.nbytes / 1e9), "episodes_per_step": E, "images_per_step": n_img,
                       "l2": "each step reads %.0f MB of crops and writes %.0f MB: beyond the 126 MB L2"
                             % ((alg_bytes[0] - n_img * 3 * IMAGE * IMAGE * 4) / 1e6, n_img * 3 * IMAGE * IMAGE * 4 / 1e6)},
            "gpu_launches": launches,
            "e2e": {"value": E * K / (ms_e2e / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": h2d // K,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K,
                    "note": "host draws class / image ids and augmentation parameters (44 B per image), uploads them, "
                            "one kernel assembles the packed fp32 episodes in HBM"},
            "clocks": clocks,
            "roofline": {"kernel": "episode_transform_kernel", "bound": "hbm", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": sum(alg_bytes) / K, "ms_per_launch": ms / K,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6500 GB/s"}}
    if not args.no_cpu_baseline:
        n_s = 2 * N_WAY * (N_SUPPORT + N_QUERY)
        d = draws[0]
        desc = store.desc_host
        host_bytes = store.data.cpu().numpy()
        imgs = []
        for p in d["params"][:n_s]:
            o, h, w = desc[int(p[0])]
            imgs.append(host_bytes[o:o + h * w * 3].reshape(h, w, 3))
        kind, dt = feeder_cpu_baseline(imgs, d["params"][:n_s], d["factors"][:n_s], IMAGE)
        line["cpu_baseline"] = {"value": 2.0 / dt, "unit": "episodes/s", "cores": 1, "kind": kind,
                                "sample": "2 episodes (210 images) of the same store and augmentation parameters through "
                                          "PIL + torchvision on one host core (the reference spreads this over 12 "
                                          "DataLoader workers, data/datamgr.py:81)"}
    print(json.dumps(line))


def tracked_traffic(key, units):
    """Bytes per launch from profiles/traffic.json: {key: {"bytes": dram read + write of the captured launch, "units": how
    many images / systems that launch processed, "source": the tracked ncu summary}} scaled to `units`; None if absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return float(t["bytes"]) / float(t["units"]) * units
    except Exception:
        return None


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _time_launch(fn, reps=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def resnet_roofline(model, lib, dev, E):
    """ResNet configs: the dominant kernel by time is conv_tcg_kernel<128> (the stride-1 3x3 layers with >= 128 channels;
    profiles/r02_launches_cfg4.summary.txt); its most frequent launch -- 3x3, 128 -> 128 at 28x28 (ResNet18 layer2; every
    stride-1 3x3 layer of the network has the same MAC count) -- is timed alone here."""
    import torch
    peaks = _peaks()
    B = E * N_WAY * (N_SUPPORT + N_QUERY)
    H = W = 28
    C = 128
    xp = torch.zeros(B, H + 2, W + 2, C, device=dev)
    xp[:, 1:-1, 1:-1].normal_()
    yp = torch.zeros(B, H + 2, W + 2, C, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    n = lib.conv_tcg_weight_floats(C, C, 3)
    wf, wd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream(dev).cuda_stream
    lib.prep_weights_tcg(w, wf, wd, C, C, 3, st)
    ms = _time_launch(lambda: lib.conv_tcg(xp, wf, None, yp, err, B, H, W, C, C, 3, st))
    assert int(err) == 0
    flops = 2.0 * B * H * W * C * C * 9
    ach = flops / (ms / 1e3) / 1e12
    peak = peaks.get("bf16_tflops", 1600.0) / 2.0
    alg_bytes = B * ((H + 2) * (W + 2) + H * W) * C * 4.0
    return {"kernel": "conv_tcg_kernel<128> (tcgen05 3xTF32) 3x3 128->128 forward, 28x28, B=%d images" % B,
            "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": tracked_traffic("conv_tcg_kernel<128>@28x28x128", B), "algorithmic_bytes_per_launch": alg_bytes,
            "algorithmic_flops_per_launch": flops, "ms_per_launch": ms,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone) / 2 = dense TF32; 3xTF32: the "
                           "arithmetic's ceiling is 1/3",
            "note": "every convolution of the network (stem, stride-1 / stride-2 3x3, 1x1, shortcuts; forward, dgrad, wgrad) runs "
                    "on tcgen05 kernels (profiles/r02_launches_cfg4.summary.txt, profiles/r02_sass_tensor_ops.txt)"}


def gp_flops(n, c, d):
    """Gram forward + backward (4 N^2 D) and C Cholesky / solve / inverse systems forward + backward (~C N^3), SURVEY 8d."""
    return 4.0 * n * n * d + float(c) * n ** 3


def run_gp_sweep(args):
    """BASELINE configs[4]: GP-only sweep on pre-extracted features Z ~ N(0,1) [E, N, 2048], 20 one-vs-rest classes:
    Gram -> C x (Cholesky, alpha, logdet, K^-1, dLoss/dK) -> dZ, per N in 25..500.  One step = E episodes of size N
    through that chain; the JSON line's value is at N = 420 (20-way 5-shot, Q = 16), the sweep rides along."""
    import torch
    import torch.distributed as dist
    from deep_kernel_transfer_b200 import _lib
    from deep_kernel_transfer_b200.engine import GPHead, GPHeadParams, make_targets
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    C, D = 20, 2048
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import gp as ogp
        torch.set_num_threads(os.cpu_count() or 1)
        N = 420
        z = torch.randn(N, D).requires_grad_(True)
        p = ogp.default_gp_params("bncossim", C, D)
        for k in ogp.trainable_gp_names("bncossim"):
            p[k].requires_grad_(True)
        tg = -torch.ones(C, N)
        for c in range(C):
            tg[c, c * 21:(c + 1) * 21] = 1.0
        ts = []
        for i in range(max(1, min(args.steps, 3)) + 1):
            t0 = time.perf_counter()
            zn = torch.nn.functional.normalize(torch.nn.functional.batch_norm(z, None, None, training=True), dim=1)
            ogp.mll_loss("bncossim", zn, tg, p).backward()
            ts.append(time.perf_counter() - t0)
        med = sorted(ts[1:])[len(ts[1:]) // 2]
        cb = {"value": 1.0 / med, "unit": "episodes/s", "cores": os.cpu_count(), "kind": "port",
              "sample": "one N = 420, C = 20, D = 2048 bncossim episode (bn_out, normalise, -mll forward + backward) per step, median of %d" % (len(ts) - 1)}
        print(json.dumps({"metric": "episodes/sec (exact-GP head only, N = 420)", "value": cb["value"], "unit": "episodes/s",
                          "n_gpus": args.gpus, "steps": len(ts) - 1, "warmup": 1, "ms_per_step": med * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "impl": "reference", "config": {"workload": "GP-only, N = 420, C = 20, D = 2048, bncossim", "bench_config": "cfg5-sweep"},
                          "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0,
                                                      "d2h_bytes_per_step": 0}}))
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    E = args.episodes_per_gpu if args.episodes_per_gpu != 32 else 8
    W, K = max(3, args.warmup), args.steps
    HP, GH = GPHeadParams(), GPHeadParams()
    HP.raw_outputscale, HP.constant = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    HP.raw_noise = torch.full((C,), -2.2532, device=dev)
    GH.raw_outputscale, GH.constant = (torch.zeros(C, device=dev) for _ in range(2))
    # bncossim (configs.kernel_type default): bn_out over the D features + F.normalize + linear kernel (variance 1)
    HP.bn_w, HP.bn_b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    HP.bn_rm, HP.bn_rv = torch.zeros(D, device=dev), torch.ones(D, device=dev)
    GH.bn_w, GH.bn_b = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    sweep = []
    result = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sizes = (25, 50, 100, 105, 180, 250, 420, 500)
    for N in sizes:
        per = max(1, N // C)
        head = GPHead(lib, "bncossim", C, D, D, 1, dev)
        head.ensure(E, N)
        z = torch.randn(E, N, D, device=dev, generator=torch.Generator(device=dev).manual_seed(rank * 100 + N))
        tg = -torch.ones(C, N, device=dev)
        for c in range(C):
            tg[c, c * per:(c + 1) * per] = 1.0

        def step(zz=None):
            f = z if zz is None else zz
            zh = head.embed(f.view(E * N, D), HP, E, N, training=True, out=head.w["zh_train"])
            loss = head.fit(zh, tg, HP, E, N, want_grad=True, grad_scale=1.0 / E)
            head.backward(f.view(E * N, D), zh, HP, GH, E, N)
            return loss
        for _ in range(W):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.launches
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.launches - l0
        head.check()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        eps = E * world * K / (ms / 1e3)
        sweep.append({"N": N, "ms_per_step": ms / K, "episodes_per_s": eps, "systems_per_s": eps * C,
                      "tflops_algorithmic": eps * gp_flops(N, C, D) / 1e12})
        if N == 420:
            # e2e: features start in pinned host memory, H2D inside the timed region, loss read back
            host = z.cpu().pin_memory()
            zb = torch.empty_like(z)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(K):
                zb.copy_(host, non_blocking=True)
                float(step(zb).sum())
            a1.record()
            barrier()
            t2 = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            # dominant kernel alone: the tiled Cholesky / inverse / gradient kernel
            w = head.w
            need = lib.gp_large_work_floats(E, C, N)
            work = torch.empty(need, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            ms_fit = _time_launch(lambda: lib.gp_fit_large(w["gram"], 0, tg, 0, HP.raw_outputscale, HP.constant, HP.raw_noise,
                                                           w["alpha"], None, w["loss_terms"], w["info"], w["dk"], w["dhyper"],
                                                           work, 1.0 / E, 1e-6, E, C, N, st))
            result = {"ms": ms, "launches": launches, "eps": eps, "ms_e2e": float(t2[0]), "h2d": host.numel() * 4,
                      "ms_fit": ms_fit}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    N = 420
    fit_flops = float(E * C) * N ** 3              # ~C N^3 per episode (Cholesky + solves + inverse + gradient products)
    ach = fit_flops / (result["ms_fit"] / 1e3) / 1e12
    peak = peaks.get("bf16_tflops", 1600.0) / 2.0
    line = {"metric": "episodes/sec (exact-GP head only, N = 420)", "value": result["eps"], "unit": "episodes/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": result["ms"] / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: GP-only sweep on pre-extracted features Z~N(0,1) [E,N,2048], 20 classes, "
                                   "bncossim (reference default): bn_out -> normalise -> Gram -> C x (Cholesky, alpha, logdet, K^-1, dK) -> dZ -> "
                                   "normalise / bn_out backward; value at N = 420",
                       "bench_config": "cfg5-sweep", "episodes_per_gpu_per_step": E, "classes": C, "D": D,
                       "l2": "per step E*C = %d systems x 2 N^2 floats of workspace + dk [E,C,N,N] = %.0f MB > 126 MB L2"
                             % (E * C, E * C * 3 * N * N * 4 / 1e6)},
            "gpu_launches": result["launches"], "tflops_algorithmic": result["eps"] * gp_flops(N, C, D) / 1e12,
            "sweep": sweep,
            "e2e": {"value": E * world * K / (result["ms_e2e"] / 1e3), "unit": "episodes/s",
                    "h2d_bytes_per_step": result["h2d"], "d2h_bytes_per_step": 4, "ms_per_step": result["ms_e2e"] / K},
            "roofline": {"kernel": "gp_fit_large_kernel (N = 420, %d systems per launch)" % (E * C), "bound": "tensor",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": tracked_traffic("gp_fit_large_kernel@N420", E * C),
                         "algorithmic_flops_per_launch": fit_flops, "ms_per_launch": result["ms_fit"],
                         "algorithmic_bytes_per_launch": float(E * C) * (2 * N * N * 4),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst) / 2 = dense TF32",
                         "note": "latency-bound factorisation (sequential pivots, barriers, L2 round trips): the tensor pipe "
                                 "is not the limiter (profiles/r01_gp_fit_large.summary.txt)"}}
    if not args.no_cpu_baseline:
        from oracle import gp as ogp
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        z1 = torch.randn(N, D).requires_grad_(True)
        p = ogp.default_gp_params("bncossim", C, D)
        for k in ogp.trainable_gp_names("bncossim"):
            p[k].requires_grad_(True)
        tg1 = -torch.ones(C, N)
        for c in range(C):
            tg1[c, c * 21:(c + 1) * 21] = 1.0

        def cpu_step():
            zn = torch.nn.functional.normalize(torch.nn.functional.batch_norm(z1, None, None, training=True), dim=1)
            ogp.mll_loss("bncossim", zn, tg1, p).backward()
        cpu_step()
        t0 = time.perf_counter()
        cpu_step()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "one N = 420, C = 20, D = 2048 bncossim episode (bn_out, normalise, -mll forward + backward) after one warm-up"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_test_path(args):
    """The test path (DKT.correct / test_loop, methods/DKT.py:199-295) at the reference's evaluation shape: 5-way 5-shot,
    15 queries per class (test.py:65-80), Conv4 bncossim, 600 episodes per test_loop.  One step = one packed call of 25
    episodes; value = episodes/s with inputs resident, e2e = test_loop() over host episodes (H2D + the hit counts back)."""
    import torch
    from deep_kernel_transfer_b200 import _lib
    from oracle import episode as oep
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    select_config("cfg3")
    n_query = 15
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        o = oep.OracleDKT("Conv4", "bncossim", n_way=5, n_support=5, seed=0)
        ts = []
        for i in range(max(1, min(args.steps, 5)) + 1):
            x = oep.synthetic_episode(i, 5, 5, n_query, 84)
            t0 = time.perf_counter()
            o.correct(x)
            ts.append(time.perf_counter() - t0)
        med = sorted(ts[1:])[len(ts[1:]) // 2]
        cb = {"value": 1.0 / med, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": "port",
              "sample": "%d single test episodes through the oracle's correct() (median)" % (len(ts) - 1)}
        print(json.dumps({"metric": "episodes/sec (meta-test)", "value": cb["value"], "unit": "episodes/s", "n_gpus": args.gpus,
                          "steps": len(ts) - 1, "warmup": 1, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                          "config": {"workload": "test path: 5-way 5-shot, 15 queries, Conv4 bncossim, 84x84", "bench_config": "test"},
                          "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0,
                                                      "d2h_bytes_per_step": 0}}))
        return
    if rank != 0:
        return      # the test path is single-process in the reference; replicas would only repeat it
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    torch.manual_seed(0)
    model = make_model(1, dev)
    model.eval()
    Et = 25
    W, K = max(3, args.warmup), args.steps
    pool = [oep.synthetic_episode(2000 + i, 5, 5, n_query, 84) for i in range(Et)]
    xs_dev = torch.stack(pool).to(dev)
    for _ in range(W):
        model.correct_packed(xs_dev)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    e0.record()
    for _ in range(K):
        hits = model.correct_packed(xs_dev)
    e1.record()
    torch.cuda.synchronize()
    t_w1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.launches - l0
    clocks = sampler.stop(t_w0, t_w1)
    # single-episode calls (the reference's granularity: one correct(x) per episode, a host sync each)
    x1 = pool[0].to(dev)
    for _ in range(3):
        model.correct(x1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        model.correct(x1)
    torch.cuda.synchronize()
    single = 20.0 / (time.perf_counter() - t0)
    # e2e: the public test_loop over host episodes (600 like test.py:65 when steps allow; at least 4 packs)
    n_ep = min(600, max(4, K) * Et)
    loader = [(pool[i % Et], None) for i in range(n_ep)]
    model.test_episodes_per_call = Et
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model.test_loop(loader[:2 * Et])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        acc = model.test_loop(loader)
        a1.record()
        torch.cuda.synchronize()
    ms_e2e = a0.elapsed_time(a1)
    n_img = 5 * (5 + n_query)
    flops = 2.0 * M_BB * n_img
    line = {"metric": "episodes/sec (meta-test)", "value": Et * K / (ms / 1e3), "unit": "episodes/s", "n_gpus": 1, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "test path (DKT.correct / test_loop): 5-way 5-shot, 15 queries per class (M = 75), Conv4 "
                                   "bncossim, 84x84; 25 episodes packed per call",
                       "bench_config": "test", "episodes_per_call": Et, "replicas": "replicas only: rank 0 measures",
                       "l2": "25 episodes = 2500 images = 212 MB of input + activations beyond the 126 MB L2"},
            "gpu_launches": launches, "tflops_algorithmic": Et * K / (ms / 1e3) * flops / 1e12,
            "single_episode_calls_per_s": single, "test_accuracy": acc,
            "e2e": {"value": n_ep / (ms_e2e / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": Et * n_img * 3 * 84 * 84 * 4,
                    "d2h_bytes_per_step": Et * 8, "ms_per_step": ms_e2e / (n_ep / Et), "episodes": n_ep,
                    "note": "DKT.test_loop over host episodes: stack + H2D of each pack, kernels, hit counts back"},
            "clocks": clocks, "roofline": dominant_kernel_roofline(model, lib, dev, Et, eval_mode=True)}
    if not args.no_cpu_baseline:
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        o = oep.OracleDKT("Conv4", "bncossim", n_way=5, n_support=5, seed=0)
        o.correct(pool[0])
        t0 = time.perf_counter()
        for i in range(3):
            o.correct(pool[i])
        dt = (time.perf_counter() - t0) / 3
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "3 single test episodes through the oracle's correct() after one warm-up"}
    print(json.dumps(line))


def dominant_kernel_roofline(model, lib, dev, E, eval_mode=False):
    """Time the layer-2 convolution forward (64->64, 42x42: 67% of the backbone MACs) alone on its stream."""
    import torch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    eng = model.feature._engine
    ws = eng.ws
    B = ws["act"][0].shape[0]
    H = W = eng.layers[1]["H"]
    st = torch.cuda.current_stream(dev).cuda_stream
    use_tc = bool(getattr(eng, "use_tc", False))

    def launch():
        eng.conv64(ws["act"][0], ws["wt_f"][1], model._P.conv_b[1], ws["y"][1], ws["partials"][1], B, H, W, st)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * H * W * 64 * 576
    achieved = flops / (ms / 1e3) / 1e12
    if use_tc:
        peak = peaks.get("bf16_tflops", 1600.0) / 2.0
        which = ("MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone) / 2 = dense TF32; 3xTF32 "
                 "error-compensated split: three tf32 products per fp32 product, i.e. the arithmetic's ceiling is 1/3")
    else:
        peak = 75.0
        which = "nominal fp32 FFMA peak (148 SMs x 128 FMA x ~1.97 GHz); no measured fp32 figure in MEASURED_PEAKS.json"
    # DRAM traffic of this kernel: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the SAME
    # kernel at a recorded image count (profiles/traffic.json, written from the tracked summary it names), scaled to this
    # launch's B; algorithmic bytes are the padded activation tensor read once + the interior of y written once
    traffic = tracked_traffic("conv3x3_tc_persistent_kernel@42x42", B) if use_tc else None
    alg_bytes = B * ((H + 2) * (W + 2) * 256.0 + H * W * 256.0)
    return {"kernel": "conv3x3 64->64 forward, 42x42, B=%d images (%s)" % (B, "tcgen05 3xTF32, conv3x3_tc_persistent_kernel" if use_tc else "fp32 FFMA"),
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms, "peak_source": which,
            "algorithmic_flops_per_launch": flops,
            "note": "per 8 input channels: a_hi x [w_hi|w_lo] (N=128) + a_lo x w_hi (N=64); the A-from-TMEM MMA forms issue at "
                            "~80 / ~60 cycles (profiles/r01_umma_microbench.log), which caps this schedule near 0.29 of dense "
                            "TF32 -- with staging and weight streaming switched off the kernel still needs 2.09 ms "
                            "(profiles/r01_tc3_experiment.txt)"}


if __name__ == "__main__":
    main()
