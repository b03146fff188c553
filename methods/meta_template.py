"""Drop-in for the reference's ``methods/meta_template.py`` import path."""
from deep_kernel_transfer_b200.methods.meta_template import MetaTemplate  # noqa: F401
