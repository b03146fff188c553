"""BASELINE configs[4]: exact-GP (Gram -> Cholesky -> MLL + gradients) scaling with the system size N, 20 one-vs-rest
classes, D = 2048 features (ResNet50), E episodes per launch; prints ms per launch and systems/s.  GPU only."""
import sys
import torch

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
C, D = 20, 2048
print("N      path     E   ms/launch  systems/s   Cholesky GFLOP/s (N^3/3 per system)")
# usage: gp_sweep.py [N ...] [--large-from N0] [--episodes E]: sizes to run; use the global-workspace kernel from N0 on
args = sys.argv[1:]
large_from = int(args[args.index("--large-from") + 1]) if "--large-from" in args else lib.gp_max_n() + 1
episodes = int(args[args.index("--episodes") + 1]) if "--episodes" in args else 8
sizes = [int(a) for i, a in enumerate(args) if a.isdigit() and (i == 0 or not args[i - 1].startswith("--"))]
for N in (sizes or (25, 50, 100, 105, 165, 180, 250, 420, 500)):
    E = episodes
    per = max(1, N // C)
    z = torch.nn.functional.normalize(torch.randn(E, N, D, device=dev), dim=2)
    targets = -torch.ones(C, N, device=dev)
    for c in range(C):
        targets[c, c * per:(c + 1) * per] = 1.0
    gram = torch.empty(E, N, N, device=dev)
    alpha = torch.empty(E, C, N, device=dev)
    lt = torch.empty(E, C, device=dev)
    info = torch.zeros(E, C, device=dev, dtype=torch.int32)
    dk = torch.empty(E, C, N, N, device=dev)
    dh = torch.empty(E, C, 3, device=dev)
    ros, cst = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    rn = torch.full((C,), -2.2532, device=dev)
    large = N >= large_from
    work = torch.empty(lib.gp_large_work_floats(E, C, N), device=dev) if large else None

    def run(with_gram=True):
        if with_gram:
            lib.gram(z, z, gram, E, N, N, D, 0)
        if large:
            lib.gp_fit_large(gram, 0, targets, 0, ros, cst, rn, alpha, None, lt, info, dk, dh, work, 1.0 / E, 0.0, E, C, N, 0)
        else:
            lib.gp_fit(gram, 0, targets, 0, ros, cst, rn, alpha, None, lt, info, dk, dh, 1.0 / E, 0.0, E, C, N, 0)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    a.record()
    for _ in range(10):
        run(False)
    b.record()
    torch.cuda.synchronize()
    ms_fit = a.elapsed_time(b) / 10
    assert int(info.abs().sum()) == 0
    print("%-6d %-8s %-3d %9.3f  %10.0f  %8.1f   fit alone %.3f ms" % (N, "global" if large else "smem", E, ms, E * C / ms * 1e3,
                                                                      E * C * N ** 3 / 3 / ms / 1e6, ms_fit))
