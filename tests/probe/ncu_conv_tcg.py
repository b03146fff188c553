"""One launch each of the ResNet tcgen05 kernels at the cfg4 batch (E = 4 episodes x 105 images = 420), for ncu:
  ncu --set full --clock-control none --import-source on -k regex:"conv_tcg_kernel|stem_tc_kernel" -c 3 -o gpurun_out/X \
      python tests/probe/ncu_conv_tcg.py
Launch order: conv_tcg_kernel<128> 3x3 128->128 at 28x28 (bench.py's roofline object), conv_tcg_kernel<64> 3x3 64->64 at
56x56, stem_tc_kernel 224x224."""
import sys

import torch

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 420
st = torch.cuda.current_stream(dev).cuda_stream
err = torch.zeros(1, device=dev, dtype=torch.int32)
for H, C in ((28, 128), (56, 64)):
    x = torch.zeros(B, H + 2, H + 2, C, device=dev)
    x[:, 1:-1, 1:-1].normal_()
    y = torch.zeros(B, H + 2, H + 2, C, device=dev)
    w = torch.randn(C, C, 3, 3, device=dev) * 0.05
    wf = torch.empty(lib.conv_tcg_weight_floats(C, C, 3), device=dev)
    lib.prep_weights_tcg(w, wf, None, C, C, 3, st)
    lib.conv_tcg(x, wf, None, y, err, B, H, H, C, C, 3, st)
xs = torch.randn(B, 3, 224, 224, device=dev)
ws = torch.randn(64, 3, 7, 7, device=dev) * 0.08
wb = torch.empty(lib.stem_tc_weight_floats(), device=dev)
lib.prep_weights_stem_tc(ws, wb, st)
ys = torch.empty(B, 112, 112, 64, device=dev)
lib.stem_tc(xs, wb, None, ys, err, B, 224, 224, st)
torch.cuda.synchronize()
assert int(err) == 0
print("ok")
