"""Per-layer timing of dktb_conv_tcg (forward; dgrad is the same kernel) at the cfg4 batch (E = 4 episodes x 105 images):
  python tests/probe/time_conv_tcg.py
Prints ms and algorithmic TFLOP/s per shape.  Cout = 64 * odd keeps the 64-channel work item, Cout % 128 == 0 takes the
128-channel one, so e.g. 128 -> 192 against 128 -> 256 compares the two per flop."""
import sys

import torch

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 420
SHAPES = [(56, 64, 64, 3), (28, 128, 192, 3), (28, 128, 128, 3), (28, 128, 256, 3), (14, 256, 192, 3), (14, 256, 256, 3),
          (7, 512, 512, 3), (7, 512, 192, 3), (56, 64, 256, 1), (56, 256, 64, 1), (28, 512, 128, 1), (14, 1024, 256, 1),
          (14, 256, 1024, 1), (7, 2048, 512, 1), (7, 512, 2048, 1)]
flush = torch.empty(64 << 20, device=dev)
for H, Cin, Cout, R in SHAPES:
    W = H
    if R == 3:
        x = torch.zeros(B, H + 2, W + 2, Cin, device=dev)
        x[:, 1:-1, 1:-1].normal_()
        y = torch.zeros(B, H + 2, W + 2, Cout, device=dev)
    else:
        x = torch.randn(B, H, W, Cin, device=dev)
        y = torch.zeros(B, H, W, Cout, device=dev)
    w = torch.randn(Cout, Cin, R, R, device=dev) * 0.05
    n = lib.conv_tcg_weight_floats(Cin, Cout, R)
    wf = torch.empty(n, device=dev)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream(dev).cuda_stream
    lib.prep_weights_tcg(w, wf, None, Cout, Cin, R, st)
    for _ in range(3):
        lib.conv_tcg(x, wf, None, y, err, B, H, W, Cin, Cout, R, st)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.conv_tcg(x, wf, None, y, err, B, H, W, Cin, Cout, R, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    assert int(err) == 0
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * B * H * W * Cin * Cout * R * R
    print("%dx%d %4d -> %4d  %dx%d  %7.3f ms  %6.1f TFLOP/s" % (H, W, Cin, Cout, R, R, ms, fl / ms / 1e9), flush=True)
