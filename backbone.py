"""Drop-in for the reference's top-level ``backbone`` module (see deep_kernel_transfer_b200/backbone.py)."""
from deep_kernel_transfer_b200.backbone import *  # noqa: F401,F403
from deep_kernel_transfer_b200.backbone import (Conv4, Conv6, Conv4NP, Conv6NP, Conv4S, Conv4SNP, ResNet10, ResNet18,  # noqa: F401
                                                ResNet34, ResNet50, ResNet101, Conv3, ConvBlock, ConvNet, Flatten,
                                                SimpleBlock, BottleneckBlock, ResNet, init_layer)
