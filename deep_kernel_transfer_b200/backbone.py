"""Feature extractors with the reference's ``backbone.py`` API surface (factory names used by
``io_utils.model_dict``, ``.trunk``, ``.final_feat_dim``, ``forward(x)``; reference backbone.py:250-268,
404-435) whose arithmetic runs in the hand-written sm_100a kernels (csrc/) instead of cuDNN.

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects below are PARAMETER CONTAINERS only: they keep the
reference's ``state_dict`` key names (``trunk.0.C.weight``, ``trunk.0.BN.running_mean``, the aliased
``trunk.0.trunk.0.weight`` ...) so checkpoints stay interchangeable; their torch ``forward`` is never
called.  The hot path (methods/DKT.py) drives ``ConvNetEngine`` directly over packed episodes;
``ConvNet.forward`` is the public single-batch entry (one BatchNorm batch = the whole input, like the
reference).  There is no CPU path: tensors must live on a CUDA device.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .engine import ConvNetEngine, ConvNetParams


def init_layer(L):
    """backbone.py:13-20: conv weights ~ N(0, sqrt(2/(k*k*out))); BN weight 1 / bias 0."""
    if isinstance(L, nn.Conv2d):
        n = L.kernel_size[0] * L.kernel_size[1] * L.out_channels
        L.weight.data.normal_(0, math.sqrt(2.0 / float(n)))
    elif isinstance(L, nn.BatchNorm2d):
        L.weight.data.fill_(1)
        L.bias.data.fill_(0)


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class ConvBlock(nn.Module):
    """Conv3x3(pad 1, bias) -> BatchNorm2d -> ReLU -> [MaxPool2d(2)]  (backbone.py:105-132)."""
    maml = False

    def __init__(self, indim, outdim, pool=True, padding=1):
        super().__init__()
        if self.maml:
            raise NotImplementedError("MAML fast-weight layers are outside the DKT hot path")
        if padding != 1:
            raise NotImplementedError("only padding=1 ConvBlocks are on the DKT hot path")
        self.indim, self.outdim, self.pool_flag = indim, outdim, pool
        self.C = nn.Conv2d(indim, outdim, 3, padding=padding)
        self.BN = nn.BatchNorm2d(outdim)
        self.relu = nn.ReLU(inplace=True)
        self.parametrized_layers = [self.C, self.BN, self.relu]
        if pool:
            self.pool = nn.MaxPool2d(2)
            self.parametrized_layers.append(self.pool)
        for layer in self.parametrized_layers:
            init_layer(layer)
        self.trunk = nn.Sequential(*self.parametrized_layers)   # aliases C/BN exactly like the reference

    def forward(self, x):
        raise RuntimeError("ConvBlock is evaluated by the fused dktb200 kernels through ConvNet.forward / DKT")


class ConvNet(nn.Module):
    def __init__(self, depth, flatten=True, image_size=84):
        super().__init__()
        if not flatten:
            raise NotImplementedError("un-flattened ConvNet is only used by RelationNet (out of scope)")
        self.depth = depth
        trunk = [ConvBlock(3 if i == 0 else 64, 64, pool=(i < 4)) for i in range(depth)]
        trunk.append(Flatten())
        self.trunk = nn.Sequential(*trunk)
        # 1600 for the reference's 84x84 inputs (backbone.py:264); other sizes are an extension used by tests
        side = image_size
        for i in range(min(depth, 4)):
            side //= 2
        self.final_feat_dim = 64 * side * side
        self._engine = None

    # -- engine plumbing -------------------------------------------------------------------------
    def blocks(self):
        return [m for m in self.trunk if isinstance(m, ConvBlock)]

    def engine_params(self):
        P = ConvNetParams(self.depth)
        for i, b in enumerate(self.blocks()):
            P.conv_w[i], P.conv_b[i] = b.C.weight.data, b.C.bias.data
            P.bn_w[i], P.bn_b[i] = b.BN.weight.data, b.BN.bias.data
            P.bn_rm[i], P.bn_rv[i] = b.BN.running_mean, b.BN.running_var
        return P

    def engine(self, image_size, device, lib=None):
        dev = torch.device(device)
        if self._engine is None or self._engine.layers[0]["H"] != image_size or self._engine.dev != dev:
            self._engine = ConvNetEngine(lib or _lib.load(), self.depth, image_size, dev)
        return self._engine

    def forward(self, x):
        """x [n,3,H,W] on a CUDA device -> features [n, 1600] in the reference's NCHW-flatten order
        (through ``bn_out`` when DKT appended it to the trunk, methods/DKT.py:45-48)."""
        if x.device.type != "cuda":
            raise RuntimeError("dktb200 has no CPU path: move the input to a CUDA device")
        n = x.shape[0]
        eng = self.engine(x.shape[-1], x.device)
        feats = eng.forward(x.contiguous().float(), self.engine_params(), ipe=n, training=self.training)
        out = feats.view(n, eng.P, 64).transpose(1, 2).reshape(n, 64 * eng.P)      # NHWC-flat -> NCHW-flat
        if hasattr(self.trunk, "bn_out"):
            bn = self.trunk.bn_out
            z = torch.empty(1, n, out.shape[1], device=x.device)
            st = [torch.empty(1, out.shape[1], device=x.device) for _ in range(3)]
            eng.lib.bn1d_fwd(out.contiguous(), bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, z,
                             st[0], st[1], st[2], 1, n, out.shape[1], 64, 1, int(self.training), 1, 0.1, 1e-5,
                             torch.cuda.current_stream(x.device).cuda_stream)
            out = z.view(n, -1)
        return out


def _not_on_hot_path(name):
    def factory(*a, **k):
        raise NotImplementedError(name + " is not on the dktb200 CUDA path (Conv4 / Conv6 / Conv3 are); there is no "
                                  "CPU or cuDNN fallback")
    factory.__name__ = name
    return factory


def Conv4():
    return ConvNet(4)


def Conv6():
    return ConvNet(6)


Conv4NP = _not_on_hot_path("Conv4NP")
Conv6NP = _not_on_hot_path("Conv6NP")
Conv4S = _not_on_hot_path("Conv4S")
Conv4SNP = _not_on_hot_path("Conv4SNP")


class SimpleBlock(nn.Module):
    """3x3 -> BN -> ReLU -> 3x3 -> BN (+ 1x1/BN shortcut) -> add -> ReLU  (backbone.py:135-185).  Parameter container."""
    maml = False
    kind = "simple"

    def __init__(self, indim, outdim, half_res):
        super().__init__()
        self.indim, self.outdim, self.half_res = indim, outdim, half_res
        self.C1 = nn.Conv2d(indim, outdim, kernel_size=3, stride=2 if half_res else 1, padding=1, bias=False)
        self.BN1 = nn.BatchNorm2d(outdim)
        self.C2 = nn.Conv2d(outdim, outdim, kernel_size=3, padding=1, bias=False)
        self.BN2 = nn.BatchNorm2d(outdim)
        self.relu1, self.relu2 = nn.ReLU(inplace=True), nn.ReLU(inplace=True)
        self.parametrized_layers = [self.C1, self.C2, self.BN1, self.BN2]
        if indim != outdim:
            self.shortcut = nn.Conv2d(indim, outdim, 1, 2 if half_res else 1, bias=False)
            self.BNshortcut = nn.BatchNorm2d(outdim)
            self.parametrized_layers += [self.shortcut, self.BNshortcut]
            self.shortcut_type = '1x1'
        else:
            self.shortcut_type = 'identity'
        for layer in self.parametrized_layers:
            init_layer(layer)


class BottleneckBlock(nn.Module):
    """1x1 -> 3x3 (with bias) -> 1x1 bottleneck; the 1x1 shortcut has NO BatchNorm (backbone.py:190-247)."""
    maml = False
    kind = "bottleneck"

    def __init__(self, indim, outdim, half_res):
        super().__init__()
        b = int(outdim / 4)
        self.indim, self.outdim, self.half_res = indim, outdim, half_res
        self.C1 = nn.Conv2d(indim, b, kernel_size=1, bias=False)
        self.BN1 = nn.BatchNorm2d(b)
        self.C2 = nn.Conv2d(b, b, kernel_size=3, stride=2 if half_res else 1, padding=1)
        self.BN2 = nn.BatchNorm2d(b)
        self.C3 = nn.Conv2d(b, outdim, kernel_size=1, bias=False)
        self.BN3 = nn.BatchNorm2d(outdim)
        self.relu = nn.ReLU()
        self.parametrized_layers = [self.C1, self.BN1, self.C2, self.BN2, self.C3, self.BN3]
        if indim != outdim:
            self.shortcut = nn.Conv2d(indim, outdim, 1, stride=2 if half_res else 1, bias=False)
            self.parametrized_layers.append(self.shortcut)
            self.shortcut_type = '1x1'
        else:
            self.shortcut_type = 'identity'
        for layer in self.parametrized_layers:
            init_layer(layer)


class ResNet(nn.Module):
    """Stem 7x7/2 -> BN -> ReLU -> MaxPool(3,2,1) -> 4 stages -> AvgPool(7) -> flatten  (backbone.py:330-376)."""
    maml = False

    def __init__(self, block, list_of_num_layers, list_of_out_dims, flatten=True):
        super().__init__()
        assert len(list_of_num_layers) == 4, 'Can have only four stages'
        if not flatten:
            raise NotImplementedError("un-flattened ResNet is only used by RelationNet (out of scope)")
        conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        bn1 = nn.BatchNorm2d(64)
        init_layer(conv1)
        init_layer(bn1)
        trunk = [conv1, bn1, nn.ReLU(), nn.MaxPool2d(kernel_size=3, stride=2, padding=1)]
        indim = 64
        for i in range(4):
            for j in range(list_of_num_layers[i]):
                half_res = (i >= 1) and (j == 0)
                trunk.append(block(indim, list_of_out_dims[i], half_res))
                indim = list_of_out_dims[i]
        trunk += [nn.AvgPool2d(7), Flatten()]
        self.final_feat_dim = indim
        self.trunk = nn.Sequential(*trunk)
        self._engine = None

    def blocks(self):
        return [m for m in self.trunk if isinstance(m, (SimpleBlock, BottleneckBlock))]

    def engine(self, image_size, device, lib=None):
        from .resnet_engine import ResNetEngine
        dev = torch.device(device)
        if self._engine is None or self._engine.dev != dev:
            self._engine = ResNetEngine(lib or _lib.load(), self, dev)
        return self._engine

    def forward(self, x):
        if x.device.type != "cuda":
            raise RuntimeError("dktb200 has no CPU path: move the input to a CUDA device")
        eng = self.engine(x.shape[-1], x.device)
        out = eng.forward(x.contiguous().float(), ipe=x.shape[0], training=self.training)
        eng.tape = []
        if hasattr(self.trunk, "bn_out"):
            bn = self.trunk.bn_out
            n, d = out.shape
            z = torch.empty(1, n, d, device=x.device)
            stt = [torch.empty(1, d, device=x.device) for _ in range(3)]
            eng.lib.bn1d_fwd(out.contiguous(), bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, z, stt[0],
                             stt[1], stt[2], 1, n, d, d, 1, int(self.training), 1, 0.1, 1e-5,
                             torch.cuda.current_stream(x.device).cuda_stream)
            out = z.view(n, d)
        return out


def ResNet10(flatten=True):
    return ResNet(SimpleBlock, [1, 1, 1, 1], [64, 128, 256, 512], flatten)


def ResNet18(flatten=True):
    return ResNet(SimpleBlock, [2, 2, 2, 2], [64, 128, 256, 512], flatten)


def ResNet34(flatten=True):
    return ResNet(SimpleBlock, [3, 4, 6, 3], [64, 128, 256, 512], flatten)


def ResNet50(flatten=True):
    return ResNet(BottleneckBlock, [3, 4, 6, 3], [256, 512, 1024, 2048], flatten)


def ResNet101(flatten=True):
    return ResNet(BottleneckBlock, [3, 4, 23, 3], [256, 512, 1024, 2048], flatten)

class Conv3(nn.Module):
    """Backbone of the QMUL head-pose regression (reference backbone.py:379-402).  Parameter containers + engine."""

    def __init__(self):
        super().__init__()
        self.layer1 = nn.Conv2d(3, 36, 3, stride=2, dilation=2)
        self.layer2 = nn.Conv2d(36, 36, 3, stride=2, dilation=2)
        self.layer3 = nn.Conv2d(36, 36, 3, stride=2, dilation=2)
        self._engine = None

    def return_clones(self):
        return [l.weight.data.clone().detach() for l in (self.layer1, self.layer2, self.layer3)]

    def assign_clones(self, weights_list):
        for l, w in zip((self.layer1, self.layer2, self.layer3), weights_list):
            l.weight.data.copy_(w)

    def layers(self):
        return (self.layer1, self.layer2, self.layer3)

    def engine(self, device, lib=None):
        from .engine import Conv3Engine
        dev = torch.device(device)
        if self._engine is None or self._engine.dev != dev:
            self._engine = Conv3Engine(lib or _lib.load(), dev)
        return self._engine

    def forward(self, x):
        """x [n,3,100,100] on a CUDA device -> [n, 2916] in the reference's NCHW-flatten order."""
        if x.device.type != "cuda":
            raise RuntimeError("dktb200 has no CPU path: move the input to a CUDA device")
        eng = self.engine(x.device)
        f = eng.forward(x.contiguous().float(), [l.weight.data for l in self.layers()], [l.bias.data for l in self.layers()])
        n = x.shape[0]
        return f.view(n, eng.P, 36).transpose(1, 2).reshape(n, -1)

