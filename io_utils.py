"""Drop-in for the part of the reference's ``io_utils`` the DKT path uses (model_dict)."""
from deep_kernel_transfer_b200.io_utils import model_dict  # noqa: F401
