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
for N in (25, 50, 100, 105, 165, 180, 250, 420, 500):
    E = 8
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
    large = N > lib.gp_max_n()
    work = torch.empty(lib.gp_large_work_floats(E, C, N), device=dev) if large else None

    def run():
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
    assert int(info.abs().sum()) == 0
    print("%-6d %-8s %-3d %9.3f  %10.0f  %8.1f" % (N, "global" if large else "smem", E, ms, E * C / ms * 1e3,
                                                 E * C * N ** 3 / 3 / ms / 1e6))
