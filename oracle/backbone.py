"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Functional CPU restatement of the
reference feature extractors (backbone.py): ConvBlock 105-132, ConvNet 250-268, SimpleBlock
135-185, BottleneckBlock 190-247, ResNet 330-376, Conv3 379-402, init_layer 13-20.

Parameters live in a flat dict keyed by the reference's ``state_dict`` names of the backbone
module (``trunk.0.C.weight`` ...), so weights exported from the reference's own classes load
directly (that is how tests/golden/make_golden.py pins this file).  PINNED by
tests/golden/backbone_*.npz.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

RESNET_CFG = {
    "ResNet10": ("simple", [1, 1, 1, 1], [64, 128, 256, 512]),
    "ResNet18": ("simple", [2, 2, 2, 2], [64, 128, 256, 512]),
    "ResNet34": ("simple", [3, 4, 6, 3], [64, 128, 256, 512]),
    "ResNet50": ("bottleneck", [3, 4, 6, 3], [256, 512, 1024, 2048]),
    "ResNet101": ("bottleneck", [3, 4, 23, 3], [256, 512, 1024, 2048]),
}


def feat_dim(arch, image_size=None):
    if arch in ("Conv4", "Conv6"):
        return 1600
    if arch == "Conv3":
        return 2916
    return RESNET_CFG[arch][2][-1]


# ----------------------------------------------------------------------------- init
def _conv_init(g, cout, cin, k, bias):
    # init_layer (backbone.py:13-20): weight ~ N(0, sqrt(2/(k*k*cout))); bias keeps torch default
    n = k * k * cout
    w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / float(n))
    out = {"weight": w}
    if bias:
        bound = 1.0 / math.sqrt(cin * k * k)
        out["bias"] = (torch.rand(cout, generator=g) * 2.0 - 1.0) * bound
    return out


def _bn_init(c):
    return {"weight": torch.ones(c), "bias": torch.zeros(c),
            "running_mean": torch.zeros(c), "running_var": torch.ones(c),
            "num_batches_tracked": torch.zeros((), dtype=torch.long)}


def _put(params, prefix, d):
    for k, v in d.items():
        params[prefix + "." + k] = v


def init_params(arch, seed=0, bn_out=False, bn_out_dim=None):
    """Random-init parameters with the reference's init_layer semantics (synthetic weights)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    if arch in ("Conv4", "Conv6"):
        depth = int(arch[-1])
        for i in range(depth):
            _put(p, "trunk.%d.C" % i, _conv_init(g, 64, 3 if i == 0 else 64, 3, True))
            _put(p, "trunk.%d.BN" % i, _bn_init(64))
    elif arch == "Conv3":
        for i, cin in enumerate((3, 36, 36)):
            bound = 1.0 / math.sqrt(cin * 9)   # Conv3 uses torch's default conv init (backbone.py:382-384)
            p["layer%d.weight" % (i + 1)] = (torch.rand(36, cin, 3, 3, generator=g) * 2 - 1) * bound
            p["layer%d.bias" % (i + 1)] = (torch.rand(36, generator=g) * 2 - 1) * bound
    else:
        kind, layers, dims = RESNET_CFG[arch]
        _put(p, "trunk.0", _conv_init(g, 64, 3, 7, False))
        _put(p, "trunk.1", _bn_init(64))
        indim, t = 64, 4
        for i in range(4):
            for j in range(layers[i]):
                outdim = dims[i]
                pre = "trunk.%d" % t
                if kind == "simple":
                    _put(p, pre + ".C1", _conv_init(g, outdim, indim, 3, False))
                    _put(p, pre + ".BN1", _bn_init(outdim))
                    _put(p, pre + ".C2", _conv_init(g, outdim, outdim, 3, False))
                    _put(p, pre + ".BN2", _bn_init(outdim))
                    if indim != outdim:
                        _put(p, pre + ".shortcut", _conv_init(g, outdim, indim, 1, False))
                        _put(p, pre + ".BNshortcut", _bn_init(outdim))
                else:
                    b = outdim // 4
                    _put(p, pre + ".C1", _conv_init(g, b, indim, 1, False))
                    _put(p, pre + ".BN1", _bn_init(b))
                    _put(p, pre + ".C2", _conv_init(g, b, b, 3, True))      # C2 has a bias (backbone.py:207)
                    _put(p, pre + ".BN2", _bn_init(b))
                    _put(p, pre + ".C3", _conv_init(g, outdim, b, 1, False))
                    _put(p, pre + ".BN3", _bn_init(outdim))
                    if indim != outdim:
                        _put(p, pre + ".shortcut", _conv_init(g, outdim, indim, 1, False))  # no BN (220-222)
                indim = outdim
                t += 1
    if bn_out:
        _put(p, "trunk.bn_out", _bn_init(bn_out_dim or feat_dim(arch)))   # DKT.py:45-48
    return p


# ----------------------------------------------------------------------------- forward
def _bn(x, p, pre, training, update):
    rm, rv = p[pre + ".running_mean"], p[pre + ".running_var"]
    if training and not update:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, p[pre + ".weight"], p[pre + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training and update and (pre + ".num_batches_tracked") in p:
        p[pre + ".num_batches_tracked"] += 1
    return y


def _conv(x, p, pre, stride=1, padding=0, dilation=1):
    return F.conv2d(x, p[pre + ".weight"], p.get(pre + ".bias"), stride, padding, dilation)


def forward(arch, p, x, training=False, update_running=True):
    """Features [n, D] in the reference's (NCHW-flatten) order; applies ``trunk.bn_out`` if present."""
    if arch in ("Conv4", "Conv6"):
        depth = int(arch[-1])
        out = x
        for i in range(depth):
            out = _conv(out, p, "trunk.%d.C" % i, padding=1)
            out = _bn(out, p, "trunk.%d.BN" % i, training, update_running)
            out = F.relu(out)
            if i < 4:
                out = F.max_pool2d(out, 2)
        out = out.reshape(out.size(0), -1)
    elif arch == "Conv3":
        out = x
        for i in (1, 2, 3):
            out = F.relu(_conv(out, p, "layer%d" % i, stride=2, dilation=2))
        out = out.reshape(out.size(0), -1)
    else:
        kind, layers, dims = RESNET_CFG[arch]
        out = _conv(x, p, "trunk.0", stride=2, padding=3)
        out = F.relu(_bn(out, p, "trunk.1", training, update_running))
        out = F.max_pool2d(out, 3, 2, 1)
        indim, t = 64, 4
        for i in range(4):
            for j in range(layers[i]):
                outdim = dims[i]
                half = (i >= 1) and (j == 0)
                s = 2 if half else 1
                pre = "trunk.%d" % t
                if kind == "simple":
                    o = _conv(out, p, pre + ".C1", stride=s, padding=1)
                    o = F.relu(_bn(o, p, pre + ".BN1", training, update_running))
                    o = _conv(o, p, pre + ".C2", padding=1)
                    o = _bn(o, p, pre + ".BN2", training, update_running)
                    if indim != outdim:
                        sh = _bn(_conv(out, p, pre + ".shortcut", stride=s), p, pre + ".BNshortcut",
                                 training, update_running)
                    else:
                        sh = out
                    out = F.relu(o + sh)
                else:
                    sh = out if indim == outdim else _conv(out, p, pre + ".shortcut", stride=s)
                    o = F.relu(_bn(_conv(out, p, pre + ".C1"), p, pre + ".BN1", training, update_running))
                    o = _conv(o, p, pre + ".C2", stride=s, padding=1)
                    o = F.relu(_bn(o, p, pre + ".BN2", training, update_running))
                    o = _bn(_conv(o, p, pre + ".C3"), p, pre + ".BN3", training, update_running)
                    out = F.relu(o + sh)
                indim = outdim
                t += 1
        out = F.avg_pool2d(out, min(7, out.shape[-1]))   # 7 at the reference's 224x224 input (backbone.py:365)
        out = out.reshape(out.size(0), -1)
    if "trunk.bn_out.weight" in p:
        out = _bn(out, p, "trunk.bn_out", training, update_running)
    return out
