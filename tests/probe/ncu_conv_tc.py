"""One training forward of the Conv4 engine at the bench shape (E = 32 episodes x 105 images = 3360 images): the first
conv3x3_tc_persistent_kernel launch is the 64 -> 64 layer at 42x42 that bench.py's roofline object describes.
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_persistent -c 1 -o gpurun_out/X python tests/probe/ncu_conv_tc.py"""
import sys

import torch

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import backbone  # noqa: E402
from deep_kernel_transfer_b200.methods.DKT import DKT  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = DKT(backbone.Conv4, 5, 5, kernel="bncossim", episodes_per_step=E).to(dev)
model.train()
model.monitor = False
x = torch.randn(E, 5, 21, 3, 84, 84, device=dev)
model._ensure_packed()
model._new_adam()
out = model.train_step(x)
torch.cuda.synchronize()
print("loss", float(out["loss"].mean()))
