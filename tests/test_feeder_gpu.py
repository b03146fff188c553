"""GPU episode feeder on the B200: bit-exact against the real PIL / torchvision golden vectors, against
the PIL-pinned oracle on seeded episodes, size-independent properties at the BASELINE episode shape, and the
``train_loop`` integration."""
import os
import sys

import numpy as np
import pytest
import torch

import kernel_checks as kc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_transforms as mg  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


@pytest.fixture(scope="module")
def lib():
    from deep_kernel_transfer_b200 import _lib
    return _lib.load()


def test_small_vs_oracle(lib):
    kc.check_episode_transform(lib, DEV)
    kc.check_episode_transform(lib, DEV, S=8, shapes=((61, 23), (30, 30), (17, 45)), seed=91, tmp_budget=1)
    kc.check_episode_transform(lib, DEV, S=7, shapes=((15, 22), (9, 9), (31, 12)), seed=92)
    kc.check_episode_transform_errors(lib, DEV)


def test_golden_pil_vectors(lib):
    """The device kernel reproduces the outputs the real PIL / torchvision gave (tests/golden/transforms.npz)."""
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore, EpisodeFeeder
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transforms.npz"))
    for size in sorted({c[3] for c in mg.AUG_CASES}):
        cases = [c for c in mg.AUG_CASES if c[3] == size]
        images = [mg.synth_image(s, h, w) for s, h, w, _ in cases]
        store = EpisodeStore(images, list(range(len(images))), DEV)
        feed = EpisodeFeeder(store, size, 1, 1, 0, aug=True, seed=0, lib=lib)
        params = np.zeros((len(cases), 8), np.int32)
        factors = np.zeros((len(cases), 3), np.float32)
        for i, (s, h, w, _) in enumerate(cases):
            top, left, ch, cw, flip = (int(v) for v in gold["aug%d_params" % s])
            params[i] = (i, top, left, ch, cw, flip, 1, 0)
            factors[i] = gold["aug%d_factors" % s]
        got = feed.transform({"params": params, "factors": factors}).cpu().numpy()
        feed.check()
        for i, (s, _, _, _) in enumerate(cases):
            assert np.array_equal(got[i], gold["aug%d_out" % s]), "aug case %d" % s
    for size in sorted({c[3] for c in mg.PLAIN_CASES}):
        cases = [c for c in mg.PLAIN_CASES if c[3] == size]
        images = [mg.synth_image(s, h, w) for s, h, w, _ in cases]
        store = EpisodeStore(images, list(range(len(images))), DEV)
        feed = EpisodeFeeder(store, size, 1, 1, 0, aug=False, seed=0, lib=lib)
        params = np.zeros((len(cases), 8), np.int32)
        for i, (s, h, w, _) in enumerate(cases):
            params[i] = (i, 0, 0, h, w, 0, 0, 0)
        got = feed.transform({"params": params, "factors": np.ones((len(cases), 3), np.float32)}).cpu().numpy()
        feed.check()
        for i, (s, _, _, _) in enumerate(cases):
            assert np.array_equal(got[i], gold["plain%d_out" % s]), "plain case %d" % s


def _cub_like_store(n_classes=8, per_class=24, seed=7):
    from deep_kernel_transfer_b200.episode_feed import EpisodeStore
    rs = np.random.RandomState(seed)
    shapes = [(375, 500), (333, 500), (500, 375), (400, 500), (281, 500), (500, 500)]
    images, labels = [], []
    for c in range(n_classes):
        for j in range(per_class):
            h, w = shapes[int(rs.randint(len(shapes)))]
            images.append(kc.synth_image(seed + 100 * c + j, h, w))
            labels.append(c)
    return images, labels, EpisodeStore(images, labels, DEV)


def test_baseline_episode_shape(lib):
    """5-way 5-shot + 16 queries at 84x84 from CUB-sized images: one episode against the oracle bit for bit, and
    properties over a packed batch (flip = mirrored output, no-jitter factors of 1 = jitter off, determinism)."""
    from deep_kernel_transfer_b200.episode_feed import EpisodeFeeder
    images, labels, store = _cub_like_store()
    feed = EpisodeFeeder(store, 84, 5, 5, 16, n_episode=4, aug=True, seed=3, lib=lib)
    d = feed.draw(4)
    x = feed.transform(d).clone()
    feed.check()
    assert x.shape == (4 * 105, 3, 84, 84) and torch.isfinite(x).all()
    one = {"params": d["params"][:105], "factors": d["factors"][:105]}
    assert np.array_equal(x[:105].cpu().numpy(), kc.feeder_oracle(images, one, 84, True))
    assert torch.equal(feed.transform(d), x)                                   # deterministic
    d2 = {"params": d["params"].copy(), "factors": d["factors"].copy()}
    d2["params"][:, 5] ^= 1
    assert torch.equal(feed.transform(d2), x.flip(-1))                         # flip is the last spatial step
    d3 = {"params": d["params"].copy(), "factors": np.ones_like(d["factors"])}
    d4 = {"params": d["params"].copy(), "factors": d["factors"]}
    d4["params"][:, 6] = 0
    assert torch.equal(feed.transform(d3).clone(), feed.transform(d4))         # factors of exactly 1 are the identity
    plain = EpisodeFeeder(store, 84, 5, 5, 16, n_episode=1, aug=False, seed=3, lib=lib)
    dp = plain.draw(1)
    xp = plain.transform(dp)
    plain.check()
    sub = {"params": dp["params"][:10], "factors": dp["factors"][:10]}
    assert np.array_equal(xp[:10].cpu().numpy(), kc.feeder_oracle(images, sub, 84, False))


def test_resnet_resolution(lib):
    """224x224 outputs (the ResNet configs): the resized image no longer leaves room for a whole-image band."""
    from deep_kernel_transfer_b200.episode_feed import EpisodeFeeder
    images, labels, store = _cub_like_store(n_classes=2, per_class=4, seed=17)
    for aug in (True, False):
        feed = EpisodeFeeder(store, 224, 2, 2, 1, n_episode=1, aug=aug, seed=5, lib=lib)
        d = feed.draw(1)
        x = feed.transform(d)
        feed.check()
        assert np.array_equal(x.cpu().numpy(), kc.feeder_oracle(images, d, 224, aug))


def test_train_loop_on_the_feeder(lib, capsys):
    """``train_loop`` consumes the feeder both ways: reference-style (x, y) items and packed device episodes."""
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    from deep_kernel_transfer_b200.episode_feed import SetDataManager
    images, labels, store = _cub_like_store(n_classes=6, per_class=8, seed=27)
    mgr = SetDataManager(84, 5, 1, 3, n_eposide=4, seed=11)
    loader = mgr.get_data_loader(store, aug=True)
    x, y = next(iter(loader))
    assert x.shape == (5, 4, 3, 84, 84) and x.is_cuda and y.shape == (5, 4)
    torch.manual_seed(0)
    model = DKT(backbone.Conv4, 5, 1, kernel="bncossim", episodes_per_step=2).cuda()
    model.train_loop(0, loader, None)
    out = capsys.readouterr().out
    assert "Epoch [0]" in out
    loss = model.last_step["loss"]
    assert loss.shape == (2,) and torch.isfinite(loss).all()
    assert loader.last_params is not None
