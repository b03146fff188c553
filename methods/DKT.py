"""Drop-in for the reference's ``methods/DKT.py`` import path (``from methods.DKT import DKT``)."""
from deep_kernel_transfer_b200.methods.DKT import DKT, kernel_type  # noqa: F401
