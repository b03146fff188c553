"""Parity tests proper: every kernel through the C ABI on the B200 against torch CPU references."""
import pytest
import torch

import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from deep_kernel_transfer_b200 import _lib
    return _lib.load()


DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def test_conv1(lib):
    kc.check_conv1(lib, DEV)
    kc.check_conv1(lib, DEV, B=4, H=84, W=84, seed=3)


def test_conv3x3_small(lib):
    kc.check_conv3x3(lib, DEV)
    kc.check_conv3x3(lib, DEV, B=2, H=10, W=10, seed=11)


def test_conv3x3_layer_shapes(lib):
    kc.check_conv3x3(lib, DEV, B=5, H=42, W=42, seed=12)
    kc.check_conv3x3(lib, DEV, B=7, H=21, W=21, seed=13)


@pytest.mark.parametrize("pool,in_pad,out_pad,H,W", [(1, 1, 1, 7, 6), (1, 0, 1, 84, 84), (0, 1, 0, 5, 5),
                                                      (1, 1, 1, 21, 21), (1, 1, 0, 10, 10)])
def test_bn_relu_pool(lib, pool, in_pad, out_pad, H, W):
    kc.check_bn_relu_pool(lib, DEV, pool=pool, in_pad=in_pad, out_pad=out_pad, H=H, W=W)


def test_head(lib):
    kc.check_head(lib, DEV)
    kc.check_head(lib, DEV, E=3, N=105, Cch=64, P=25, seed=8)


def test_gp(lib):
    kc.check_gp(lib, DEV)
    kc.check_gp(lib, DEV, E=4, C=5, per_class=21, D=1600, M=75, seed=9)     # N = 105 (5-way 5-shot, Q=16)
    kc.check_gp(lib, DEV, E=2, C=5, per_class=17, D=1600, M=75, seed=10)    # N = 85  (5-way 1-shot)
    kc.check_gp(lib, DEV, E=2, C=5, per_class=1, D=64, M=75, seed=11)       # N = 5   (1-shot test episode)


@pytest.mark.parametrize("fn", ["conv1_bwd_fused", "conv1_bwd_fused_mma"])
def test_conv1_bwd_fused(lib, fn):
    kc.check_conv1_bwd_fused(lib, DEV, fn=fn)
    kc.check_conv1_bwd_fused(lib, DEV, E=1, ipe=3, H=9, W=8, out_pad=0, seed=22, fn=fn)
    kc.check_conv1_bwd_fused(lib, DEV, E=2, ipe=5, H=84, W=84, out_pad=1, seed=23, fn=fn)   # BASELINE geometry


def test_gp_large(lib):
    """N beyond shared memory (BASELINE configs[4]: 20-way 5-shot, Gram-N sweep up to 500)."""
    kc.check_gp(lib, DEV, E=2, C=5, per_class=21, D=1600, M=75, seed=9, large=True)      # N = 105, same as small path
    kc.check_gp(lib, DEV, E=2, C=5, per_class=36, D=1600, M=80, seed=12, large=True)     # N = 180 (5-way 20-shot)
    kc.check_gp(lib, DEV, E=1, C=20, per_class=21, D=512, M=100, seed=13, large=True)   # N = 420 (20-way 5-shot)
    kc.check_gp(lib, DEV, E=1, C=4, per_class=128, D=512, M=64, seed=14, large=True)    # N = 512


def test_gp_jitter_retry(lib):
    """psd_safe_cholesky retries (1e-6, 1e-5, 1e-4), NaN loss + zero gradients on a system that stays indefinite."""
    kc.check_gp_jitter(lib, DEV)
    kc.check_gp_jitter(lib, DEV, N=105, seed=92)
    kc.check_gp_jitter(lib, DEV, N=70, large=True, seed=91)
    kc.check_gp_jitter(lib, DEV, N=420, large=True, seed=93)


def test_adam(lib):
    kc.check_adam(lib, DEV, n=100003)


@pytest.mark.parametrize("fn", ["conv3x3_tc_fwd"])
def test_conv3x3_tc(lib, fn):
    """tcgen05 + TMA 3xTF32 convolution (forward and dgrad) at small and at the real layer shapes."""
    kc.check_conv3x3_tc(lib, DEV, fn=fn)
    kc.check_conv3x3_tc(lib, DEV, B=2, H=10, W=10, seed=22, fn=fn)
    kc.check_conv3x3_tc(lib, DEV, B=4, H=42, W=42, seed=23, fn=fn)
    kc.check_conv3x3_tc(lib, DEV, B=5, H=21, W=21, seed=24, fn=fn)


def test_conv3x3_wgrad_tc(lib):
    kc.check_conv3x3_wgrad_tc(lib, DEV)
    kc.check_conv3x3_wgrad_tc(lib, DEV, B=2, H=10, W=10, seed=32)
    kc.check_conv3x3_wgrad_tc(lib, DEV, B=6, H=42, W=42, seed=33)
    kc.check_conv3x3_wgrad_tc(lib, DEV, B=64, H=21, W=21, seed=34)


@pytest.mark.parametrize("kernel", ["linear", "rbf", "matern", "poli1", "poli2"])
def test_gp_family(lib, kernel):
    kc.check_gp_family(lib, DEV, kernel)
    kc.check_gp_family(lib, DEV, kernel, E=3, C=5, per_class=21, D=512, M=75, seed=41)   # cfg4 shape (N=105, D=512)


@pytest.mark.parametrize("cfg", [dict(), dict(Cin=36, Cout=36, H=48, W=48, N=19), dict(Cin=5, Cout=70, R=1, stride=2, dil=1, relu=0),
                                 dict(Cin=3, Cout=64, R=7, stride=2, pad=3, dil=1, relu=0, H=56, W=56, N=3),
                                 dict(N=19, H=100, W=100),
                                 dict(Cin=64, Cout=64, R=3, stride=1, pad=1, dil=1, relu=0, H=56, W=56, N=5),    # ResNet18 layers
                                 dict(Cin=64, Cout=128, R=3, stride=2, pad=1, dil=1, relu=0, H=56, W=56, N=3),
                                 dict(Cin=64, Cout=128, R=1, stride=2, pad=0, dil=1, relu=0, H=56, W=56, N=3),
                                 dict(Cin=256, Cout=512, R=3, stride=2, pad=1, dil=1, relu=0, H=14, W=14, N=4),
                                 dict(Cin=512, Cout=512, R=3, stride=1, pad=1, dil=1, relu=0, H=7, W=7, N=6),
                                 dict(Cin=256, Cout=64, R=1, stride=1, pad=0, dil=1, relu=0, H=56, W=56, N=2),   # bottleneck 1x1
                                 dict(Cin=32, Cout=48, R=3, stride=1, pad=2, dil=2, relu=0, H=11, W=9, N=2),     # dilation, ragged tiles
                                 dict(Cin=96, Cout=64, R=3, stride=2, pad=1, dil=1, relu=0, H=15, W=13, N=3)])   # odd sizes through a stride
def test_conv2d_generic(lib, cfg):
    kc.check_conv2d(lib, DEV, **cfg)


def test_conv1_tc(lib):
    kc.check_conv1_tc(lib, DEV)
    kc.check_conv1_tc(lib, DEV, E=1, ipe=2, H=32, W=32, seed=71)
    kc.check_conv1_tc(lib, DEV, E=2, ipe=2, H=21, W=37, seed=72)


def test_resnet_ops(lib):
    kc.check_resnet_ops(lib, DEV)
    kc.check_resnet_ops(lib, DEV, E=2, ipe=3, H=14, W=14, C=256, seed=81)
    kc.check_resnet_ops(lib, DEV, E=2, ipe=2, H=32, W=33, C=64, seed=82)      # BatchNorm sums split over the pixels (4 splits)


@pytest.mark.parametrize("cfg", [dict(), dict(B=3, H=56, W=56, C=64, Cout=128, seed=131),             # ResNet18 layer2 down-sampling
                                 dict(B=3, H=28, W=28, C=128, Cout=256, seed=132, x_pad=True),  # layer3, padded-flat input
                                 dict(B=5, H=14, W=14, C=256, Cout=512, seed=133, bias=True),   # layer4
                                 dict(B=2, H=28, W=28, C=128, Cout=128, seed=134),              # ResNet50 bottleneck C2
                                 dict(B=1, H=6, W=122, C=64, Cout=64, seed=135)])               # widest supported row
def test_conv_tcg_s2(lib, cfg):
    kc.check_conv_tcg_s2(lib, DEV, **cfg)


@pytest.mark.parametrize("cfg", [dict(), dict(B=1, H=20, W=30, seed=121, bias=True),              # ragged band and column tile
                                 dict(B=2, H=84, W=84, seed=122),                                # miniImageNet resolution
                                 dict(B=3, H=224, W=224, seed=123)])                             # the reference's ResNet resolution
def test_stem_tc(lib, cfg):
    kc.check_stem_tc(lib, DEV, **cfg)


@pytest.mark.parametrize("cfg", [dict(), dict(B=3, H=56, W=56, Cin=64, Cout=64, seed=101),          # ResNet18 layer1
                                 dict(B=3, H=28, W=28, Cin=128, Cout=128, seed=102, bias=False),      # layer2
                                 dict(B=5, H=14, W=14, Cin=256, Cout=256, seed=103),                  # layer3
                                 dict(B=7, H=7, W=7, Cin=512, Cout=512, seed=104, bias=False),        # layer4
                                 dict(B=2, H=56, W=56, Cin=64, Cout=256, R=1, seed=105),              # bottleneck 1x1 up
                                 dict(B=2, H=28, W=28, Cin=512, Cout=128, R=1, seed=106, bias=False),  # bottleneck 1x1 down
                                 dict(B=3, H=7, W=7, Cin=2048, Cout=512, R=1, seed=107),              # ResNet50 layer4
                                 dict(B=1, H=9, W=61, Cin=192, Cout=64, seed=108)])                   # widest supported row, odd group count
def test_conv_tcg(lib, cfg):
    """tcgen05 / TMA convolution of the ResNet layers: forward + dgrad, 3x3 padded-flat and 1x1 flat."""
    kc.check_conv_tcg(lib, DEV, **cfg)


@pytest.mark.parametrize("cfg", [dict(), dict(E=2, M=420, N=420, D=2048, seed=111), dict(E=4, M=85, N=85, D=1600, seed=112),
                                 dict(E=2, M=75, N=105, D=1600, same=False, seed=113),         # prediction: queries x train
                                 dict(E=2, M=300, N=100, D=512, same=False, seed=114),
                                 dict(E=40, M=5, N=5, D=64, seed=115), dict(E=8, M=20, N=20, D=40, seed=116)])
def test_gram_tc(lib, cfg):
    kc.check_gram_tc(lib, DEV, **cfg)
