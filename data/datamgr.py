"""Drop-in for the reference's ``data/datamgr.py`` import path (``from data.datamgr import SetDataManager``,
train.py:14 / test.py:17): the episodic loader served by the GPU episode feeder."""
from deep_kernel_transfer_b200.episode_feed import SetDataManager, EpisodeStore, EpisodeFeeder  # noqa: F401
