"""Debugging aid: replays the op tape of a ResNet forward with plain torch autograd (fp64 and fp32) and compares the
engine's parameter gradients with both -- separates 'composition differs' from 'rounding differs'.  Also compares the
features with the oracle backbone (fp64).  usage: resnet_chain_check.py ARCH SIZE [emu]"""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from deep_kernel_transfer_b200 import backbone, _lib          # noqa: E402
from deep_kernel_transfer_b200.resnet_engine import ResNetEngine  # noqa: E402
from oracle import backbone as obb                            # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 224
emu = len(sys.argv) > 3 and sys.argv[3] == "emu"
if emu:
    from emu import build_emu
    lib = _lib.DktbLib(build_emu.build())
    dev = torch.device("cpu")
else:
    lib = _lib.load()
    dev = torch.device("cuda:0")
B, ipe = 4, 4
torch.manual_seed(0)
net = getattr(backbone, arch)().to(dev)
g = torch.Generator().manual_seed(3)
for n_, p in net.named_parameters():
    if p.dim() == 1:
        p.data.add_(0.1 * torch.randn(p.shape, generator=g).to(dev))
    p.grad = torch.zeros_like(p)
eng = ResNetEngine(lib, net, dev)
x = torch.randn(B, 3, size, size, generator=g).to(dev)
feats = eng.forward(x, ipe, True)
tape = list(eng.tape)
gf = torch.randn(feats.shape, generator=g).to(dev)
eng.backward(gf)
names = {id(p): n_ for n_, p in net.named_parameters()}


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max().clamp_min(1e-30))


def replay(dtype):
    vals = {}
    params = {}

    def P(t):
        if id(t) not in params:
            params[id(t)] = t.detach().to(dtype).clone().requires_grad_(True)
        return params[id(t)]

    first = tape[0]
    vals[first[1].data_ptr()] = first[1].detach().to(dtype).permute(0, 3, 1, 2)
    for rec in tape:
        if rec[0] == "conv":
            _, xi, out, m, (Bn, H, W, Cin, R, st, pad, dil) = rec
            vals[out.data_ptr()] = F.conv2d(vals[xi.data_ptr()], P(m.weight), None, st, pad, dil)
        elif rec[0] == "bn":
            _, xi, y, m, mean, invstd, res, relu, ipe_ = rec
            o = F.batch_norm(vals[xi.data_ptr()], None, None, P(m.weight), P(m.bias), True, 0.0, 1e-5)
            if res is not None:
                o = o + vals[res.data_ptr()]
            vals[y.data_ptr()] = o.relu() if relu else o
        elif rec[0] == "maxpool":
            vals[rec[2].data_ptr()] = F.max_pool2d(vals[rec[1].data_ptr()], 3, 2, 1)
        elif rec[0] == "avgpool":
            v = vals[rec[1].data_ptr()]
            vals[rec[2].data_ptr()] = F.avg_pool2d(v, v.shape[-1]).flatten(1)
    f = vals[tape[-1][2].data_ptr()]
    f.backward(gf.to(dtype))
    return f, vals, params


f64, v64, p64 = replay(torch.float64)
f32, v32, p32 = replay(torch.float32)
print("features: engine vs chain64 %.2e   chain32 vs chain64 %.2e" % (rel(feats, f64), rel(f32, f64)))
sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
of = obb.forward(arch, sd, x.double().cpu(), training=True) if hasattr(obb, "forward") else None
if of is not None:
    of = of[0] if isinstance(of, tuple) else of
    print("features: chain64 vs oracle backbone %.2e" % rel(f64.cpu(), of))
print("== activations (engine vs chain64 | chain32 vs chain64)")
for i, rec in enumerate(tape):
    out = rec[2]
    a = out if out.dim() == 2 else out.permute(0, 3, 1, 2)
    print(i, rec[0], tuple(out.shape), "%.2e | %.2e" % (rel(a, v64[out.data_ptr()]), rel(v32[out.data_ptr()], v64[out.data_ptr()])))
print("== parameter gradients (engine vs chain64 | chain32 vs chain64)")
for n_, p in net.named_parameters():
    if id(p) in p64:
        print(n_, "%.2e | %.2e" % (rel(p.grad, p64[id(p)].grad), rel(p32[id(p)].grad, p64[id(p)].grad)))
