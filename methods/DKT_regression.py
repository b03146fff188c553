"""Drop-in for the reference's ``methods/DKT_regression.py`` import path (``from methods.DKT_regression import DKT``)."""
from deep_kernel_transfer_b200.methods.DKT_regression import DKT, ExactGPLayer  # noqa: F401
