"""GPU debugging aid: every op of a ResNet forward/backward at the reference resolution is re-evaluated with plain
torch in fp64 from the engine's OWN input tensors, so an error is attributed to the op that makes it."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from deep_kernel_transfer_b200 import backbone, _lib          # noqa: E402
from deep_kernel_transfer_b200.resnet_engine import ResNetEngine  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 224
B, ipe = 4, 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = getattr(backbone, arch)().to(dev)
for p in net.parameters():
    p.grad = torch.zeros_like(p)
eng = ResNetEngine(_lib.load(), net, dev)
x = torch.randn(B, 3, size, size, device=dev)
feats = eng.forward(x, ipe, True)
tape = list(eng.tape)


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def nchw(t):
    return t.double().permute(0, 3, 1, 2)


print("== forward")
for i, rec in enumerate(tape):
    if rec[0] == "conv":
        _, xi, out, m, (Bn, H, W, Cin, R, st, pad, dil) = rec
        ref = F.conv2d(nchw(xi), m.weight.double(), None, st, pad, dil).permute(0, 2, 3, 1)
        print(i, "conv", tuple(out.shape), "%.2e" % rel(out, ref))
    elif rec[0] == "bn":
        _, xi, y, m, mean, invstd, res, relu, ipe_ = rec
        ref = F.batch_norm(nchw(xi), None, None, m.weight.double(), m.bias.double(), True, 0.0, 1e-5)
        if res is not None:
            ref = ref + nchw(res)
        if relu:
            ref = ref.relu()
        print(i, "bn", tuple(y.shape), "%.2e" % rel(y, ref.permute(0, 2, 3, 1)))
gf = torch.randn_like(feats)
eng.trace = []
eng.backward(gf)
print("== backward")
for rec, gy, gx, extra in eng.trace:
    if rec[0] == "conv":
        _, xi, out, m, (Bn, H, W, Cin, R, st, pad, dil) = rec
        xin = nchw(xi).detach().requires_grad_(True)
        w = m.weight.detach().double().requires_grad_(True)
        o = F.conv2d(xin, w, None, st, pad, dil)
        o.backward(nchw(gy))
        print("conv", tuple(out.shape), "dw %.2e" % rel(extra, w.grad), "dx %.2e" % (rel(gx, xin.grad.permute(0, 2, 3, 1)) if gx is not None else 0))
    else:
        _, xi, y, m, mean, invstd, res, relu, ipe_ = rec
        xin = nchw(xi).detach().requires_grad_(True)
        g_ = m.weight.detach().double().requires_grad_(True)
        b_ = m.bias.detach().double().requires_grad_(True)
        r_ = None if res is None else nchw(res).detach().requires_grad_(True)
        o = F.batch_norm(xin, None, None, g_, b_, True, 0.0, 1e-5)
        if r_ is not None:
            o = o + r_
        if relu:
            o = o.relu()
        o.backward(nchw(gy))
        msg = "dx %.2e dg %.2e db %.2e" % (rel(gx, xin.grad.permute(0, 2, 3, 1)), rel(m.weight.grad, g_.grad), rel(m.bias.grad, b_.grad))
        if r_ is not None:
            msg += " dres %.2e" % rel(extra, r_.grad.permute(0, 2, 3, 1))
        print("bn", tuple(y.shape), msg)
