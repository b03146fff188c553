"""Small invocations of every tcgen05 / TMA / mbarrier kernel through the C ABI, for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tests/probe/sanitize_tc.py
  compute-sanitizer --tool racecheck python tests/probe/sanitize_tc.py
(the parity checks of tests/kernel_checks.py at shapes that keep the instrumented run short)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kernel_checks as kc  # noqa: E402
from deep_kernel_transfer_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["conv3x3", "wgrad", "conv1", "tcg", "gram", "stem", "s2"]
if "conv3x3" in which:
    kc.check_conv3x3_tc(lib, dev, B=2, H=10, W=10, seed=22, fn="conv3x3_tc_fwd")
    print("conv3x3_tc ok")
if "wgrad" in which:
    kc.check_conv3x3_wgrad_tc(lib, dev, B=2, H=10, W=10, seed=32)
    print("conv3x3_wgrad_tc ok")
if "conv1" in which:
    kc.check_conv1_tc(lib, dev, E=1, ipe=2, H=32, W=32, seed=71)
    print("conv1_tc ok")
if "tcg" in which:
    kc.check_conv_tcg(lib, dev, B=1, H=6, W=5, Cin=128, Cout=64, R=3, seed=100)
    kc.check_conv_tcg(lib, dev, B=1, H=6, W=5, Cin=64, Cout=128, R=1, seed=101)
    kc.check_conv_tcg(lib, dev, B=1, H=6, W=5, Cin=128, Cout=128, R=3, seed=102)       # the 128-channel work item
    kc.check_conv_tcg(lib, dev, B=1, H=5, W=4, Cin=256, Cout=128, R=1, seed=103)
    print("conv_tcg ok")
if "gram" in which and lib.has("dktb_gram_tc"):
    kc.check_gram_tc(lib, dev, E=2, M=70, N=70, D=96, seed=111)
    print("gram_tc ok")
if "stem" in which and lib.has("dktb_stem_tc"):
    kc.check_stem_tc(lib, dev, B=1, H=20, W=30, seed=121, bias=True)
    print("stem_tc ok")
if "s2" in which and lib.has("dktb_conv_tcg_s2"):
    kc.check_conv_tcg_s2(lib, dev, B=1, H=8, W=10, C=64, Cout=128, seed=130)
    kc.check_conv_tcg_s2(lib, dev, B=1, H=6, W=8, C=128, Cout=128, seed=136, x_pad=True, bias=True)
    print("conv_tcg_s2 ok")
torch.cuda.synchronize()
print("done")
