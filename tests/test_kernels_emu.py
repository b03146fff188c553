"""Host-logic tests (no GPU): the plain-CUDA kernel sources compiled with g++ against the CPU
emulation of the CUDA execution model (tests/emu) and driven through the same C ABI + ctypes binding
as the product.  They check index math / staging / reductions of every kernel against torch on CPU.
The emulation library is test infrastructure: the product never loads it."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402
import kernel_checks as kc  # noqa: E402
from deep_kernel_transfer_b200._lib import DktbLib  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return DktbLib(build_emu.build())


DEV = torch.device("cpu")


def test_conv1(lib):
    kc.check_conv1(lib, DEV)


def test_conv3x3(lib):
    kc.check_conv3x3(lib, DEV)


def test_conv3x3_odd(lib):
    kc.check_conv3x3(lib, DEV, B=2, H=10, W=10, seed=11)


@pytest.mark.parametrize("pool,in_pad,out_pad,H,W", [(1, 1, 1, 7, 6), (1, 0, 1, 8, 8), (0, 1, 0, 5, 5), (1, 1, 0, 10, 10)])
def test_bn_relu_pool(lib, pool, in_pad, out_pad, H, W):
    kc.check_bn_relu_pool(lib, DEV, pool=pool, in_pad=in_pad, out_pad=out_pad, H=H, W=W)


@pytest.mark.parametrize("fn", ["conv1_bwd_fused", "conv1_bwd_fused_mma"])
def test_conv1_bwd_fused(lib, fn):
    kc.check_conv1_bwd_fused(lib, DEV, fn=fn)
    kc.check_conv1_bwd_fused(lib, DEV, E=1, ipe=3, H=9, W=8, out_pad=0, seed=22, fn=fn)     # odd height: dropped edge row


def test_head(lib):
    kc.check_head(lib, DEV)
    kc.check_head(lib, DEV, E=1, N=5, Cch=12, P=1, seed=8)


def test_gp(lib):
    kc.check_gp(lib, DEV)


def test_gp_large(lib):
    kc.check_gp(lib, DEV, E=1, C=2, per_class=20, D=24, M=5, seed=6, large=True)     # N = 40: one partial tile
    kc.check_gp(lib, DEV, E=1, C=3, per_class=23, D=32, M=5, seed=7, large=True)     # N = 69: two tiles
    kc.check_gp(lib, DEV, E=1, C=1, per_class=133, D=48, M=5, seed=8, large=True)    # N = 133: three tiles (5 rows in the last)


def test_gp_jitter_retry(lib):
    kc.check_gp_jitter(lib, DEV)
    kc.check_gp_jitter(lib, DEV, N=70, large=True, seed=91)


def test_adam(lib):
    kc.check_adam(lib, DEV)


@pytest.mark.parametrize("kernel", ["linear", "rbf", "matern", "poli1", "poli2"])
def test_gp_family(lib, kernel):
    kc.check_gp_family(lib, DEV, kernel)


@pytest.mark.parametrize("cfg", [dict(), dict(Cin=36, Cout=36, H=12, W=12), dict(Cin=5, Cout=70, R=1, stride=2, dil=1, relu=0),
                                 dict(Cin=3, Cout=20, R=7, stride=2, pad=3, dil=1, relu=0, H=17, W=15),
                                 dict(Cin=32, Cout=64, R=3, stride=1, pad=1, dil=1, relu=0, H=5, W=4, N=2),      # tensor-core tiles
                                 dict(Cin=64, Cout=32, R=3, stride=2, pad=1, dil=1, relu=0, H=7, W=6, N=1),
                                 dict(Cin=18, Cout=10, R=1, stride=1, pad=0, dil=1, relu=0, H=6, W=5, N=1),
                                 dict(Cin=32, Cout=12, R=3, stride=2, pad=2, dil=2, relu=0, H=7, W=5, N=1)])
def test_conv2d_generic(lib, cfg):
    kc.check_conv2d(lib, DEV, **cfg)


def test_spectral(lib):
    kc.check_spectral(lib, DEV)


def test_resnet_ops(lib):
    kc.check_resnet_ops(lib, DEV)
    kc.check_resnet_ops(lib, DEV, E=2, ipe=2, H=24, W=23, C=8, seed=82)       # BatchNorm sums split over the pixels (2 splits)


def test_episode_transform(lib):
    kc.check_episode_transform(lib, DEV)


def test_episode_transform_banded(lib):
    kc.check_episode_transform(lib, DEV, S=8, shapes=((61, 23), (30, 30), (17, 45)), seed=91, tmp_budget=1)   # several bands per image


def test_episode_transform_odd_size(lib):
    kc.check_episode_transform(lib, DEV, S=7, shapes=((15, 22), (9, 9), (31, 12)), seed=92)     # rows are not whole words


def test_episode_transform_errors(lib):
    kc.check_episode_transform_errors(lib, DEV)
