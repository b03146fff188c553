"""Drop-in for the reference's top-level ``io_utils`` (io_utils.py:7-15 model_dict, 66-86 checkpoint file helpers).
With a reference checkout behind this repo on sys.path, everything else the drivers import from it (``parse_args``)
is passed through from the reference's own file; ``model_dict`` always holds this repo's backbones."""
import glob
import os

import numpy as np

from deep_kernel_transfer_b200._compat import load_shadowed
from deep_kernel_transfer_b200.io_utils import model_dict as _model_dict

_ref = load_shadowed("io_utils.py", __file__, "_reference_io_utils")
if _ref is not None:
    for _k, _v in vars(_ref).items():
        if not _k.startswith("__"):
            globals().setdefault(_k, _v)
model_dict = _model_dict


def get_assigned_file(checkpoint_dir, num):
    return os.path.join(checkpoint_dir, "{:d}.tar".format(num))


def get_resume_file(checkpoint_dir):
    """The numbered checkpoint with the largest epoch (``best_model.tar`` excluded), or None."""
    files = [f for f in glob.glob(os.path.join(checkpoint_dir, "*.tar")) if os.path.basename(f) != "best_model.tar"]
    if not files:
        return None
    epochs = np.array([int(os.path.splitext(os.path.basename(f))[0]) for f in files])
    return os.path.join(checkpoint_dir, "{:d}.tar".format(int(epochs.max())))


def get_best_file(checkpoint_dir):
    best = os.path.join(checkpoint_dir, "best_model.tar")
    return best if os.path.isfile(best) else get_resume_file(checkpoint_dir)
