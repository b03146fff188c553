"""Drop-in for the reference's ``data/datamgr.py`` import path (``from data.datamgr import SimpleDataManager,
SetDataManager``, train.py:14 / test.py:16): the episodic loader is the GPU episode feeder; the non-episodic
``SimpleDataManager`` (baseline pre-training, out of scope) is passed through from a reference checkout when one is on
sys.path."""
from deep_kernel_transfer_b200._compat import load_shadowed
from deep_kernel_transfer_b200.episode_feed import SetDataManager, EpisodeStore, EpisodeFeeder  # noqa: F401

_ref = None
try:
    _ref = load_shadowed("data/datamgr.py", __file__, "_reference_data_datamgr")
except ImportError:          # the reference's loader needs PIL / torchvision
    _ref = None

if _ref is not None:
    SimpleDataManager = _ref.SimpleDataManager
    TransformLoader = _ref.TransformLoader
    ReferenceSetDataManager = _ref.SetDataManager
else:
    class SimpleDataManager(object):
        def __init__(self, *a, **k):
            raise NotImplementedError("SimpleDataManager (non-episodic baseline loader) is outside the DKT hot path; "
                                      "put a checkout of the reference behind this repo on sys.path to use it")
