"""BASELINE configs[3]: 5-way 5-shot ResNet18, RBF kernel, synthetic 224x224 episodes -- meta-train steps per second of
the drop-in module on one GPU (the ResNet layers run on the generic fp32 CUDA-core kernels this round).  GPU only.
usage: resnet_bench.py [arch] [episodes_per_step] [n_way]      (n_way 20 = the end-to-end point of configs[4], N = 420)"""
import sys
import time

import torch

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import backbone  # noqa: E402
from deep_kernel_transfer_b200.methods.DKT import DKT  # noqa: E402
from oracle import episode as oep  # noqa: E402  (synthetic episode generator only)

arch = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1
WAY = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
model = DKT(getattr(backbone, arch), WAY, 5, kernel="rbf", episodes_per_step=E).to(dev)
model.train()
model._ensure_packed()
model._new_adam()
xs = torch.stack([oep.synthetic_episode(e, WAY, 5, 16, 224) for e in range(E)]).to(dev)
for _ in range(2):
    out = model.train_step(xs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
a.record()
for _ in range(K):
    out = model.train_step(xs)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
macs = {"ResNet10": 0.889e9, "ResNet18": 1.814e9, "ResNet34": 3.663e9, "ResNet50": 4.087e9}[arch]
flop = E * WAY * 21 * (6 + 2) * macs          # fwd + bwd (3 x 2 flop per MAC) + monitoring forward
print("%s rbf %d-way 5-shot Q=16 @224, E=%d: %.1f ms/step, %.2f episodes/s, %.1f algorithmic TFLOP/s, loss %.4f, "
      "peak memory %.1f GB" % (arch, WAY, E, ms, E / ms * 1e3, flop / ms / 1e9, float(out["loss"].mean()),
                              torch.cuda.max_memory_allocated() / 2 ** 30))
