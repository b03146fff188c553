"""Drop-in for the reference's top-level ``configs`` module."""
from deep_kernel_transfer_b200.configs import *  # noqa: F401,F403
from deep_kernel_transfer_b200.configs import kernel_type, save_dir, data_dir  # noqa: F401
