"""GPU episode feeder: a device-resident uint8 image store plus on-device episode assembly -- the
drop-in for the reference's ``SetDataManager`` / ``SetDataset`` / ``EpisodicBatchSampler`` / ``TransformLoader`` chain
(data/datamgr.py:14-84, data/dataset.py:29-87, data/additional_transforms.py:19-34).

The reference decodes, crops, jitters and normalises every image of every episode with PIL in 12 DataLoader workers and
ships fp32 tensors to the GPU (8.9 MB per 5-way 5-shot episode).  Here the decoded images live in HBM as uint8 (CUB:
~6.6 GB, miniImagenet: ~34 GB of 180 GB); per meta-step the host only draws the episode composition and the
augmentation parameters (class ids, image ids, crop boxes, jitter factors, flips: 44 bytes per image) and ONE kernel
(``dktb_episode_transform``, csrc/episode_feed.cu) writes the packed fp32 ``[E, C, S+Q, 3, H, W]`` tensor, bit-identical to
what PIL / torchvision produce for the same parameters.

Sampling follows the reference: ``torch.randperm(n_classes)[:n_way]`` per episode (dataset.py:86), the first S+Q entries
of a fresh ``torch.randperm(len(class))`` per sampled class (the shuffling sub-DataLoader of dataset.py:47-56), and per
image torchvision's RandomResizedCrop box search, ``0.4 * (2 u - 1) + 1`` jitter factors and a fair flip coin -- drawn in
vectorised form from one ``torch.Generator`` (the reference's random STREAM depends on its worker processes and is not
reproducible; the distributions are the contract).
"""
import json
import math

import numpy as np
import torch

from . import _lib

MEAN = (0.485, 0.456, 0.406)      # data/datamgr.py:16
STD = (0.229, 0.224, 0.225)
JITTER = (0.4, 0.4, 0.4)          # Brightness, Contrast, Color (data/datamgr.py:17)


class EpisodeStore:
    """Decoded RGB images, HWC uint8, back to back in one device buffer (each image 16-byte aligned)."""

    def __init__(self, images, labels, device):
        if len(images) != len(labels) or len(images) == 0:
            raise ValueError("EpisodeStore needs one label per image and at least one image")
        self.device = torch.device(device)
        desc = np.zeros((len(images), 3), np.int64)
        off = 0
        for i, im in enumerate(images):
            h, w, c = im.shape
            if c != 3:
                raise ValueError("image %d is not HWC RGB" % i)
            desc[i] = (off, h, w)
            off += (h * w * 3 + 15) & ~15
        host = torch.empty(off, dtype=torch.uint8)
        hv = host.numpy()
        for i, im in enumerate(images):
            o, h, w = desc[i]
            hv[o:o + h * w * 3] = np.asarray(im, dtype=np.uint8).reshape(-1)
        self.data = host.to(self.device)
        self.desc_host = desc
        self.desc = torch.from_numpy(desc).to(self.device)
        self.labels = np.asarray(labels)
        self.cl_list = np.unique(self.labels).tolist()               # dataset.py:33
        self.sub_meta = [np.nonzero(self.labels == cl)[0] for cl in self.cl_list]     # dataset.py:35-40
        self.nbytes = off

    def __len__(self):
        return self.desc_host.shape[0]

    @classmethod
    def from_device_bytes(cls, shapes, labels, device, seed=0):
        """A store of random bytes generated on the device (benchmarks: the content does not change the work)."""
        self = cls.__new__(cls)
        self.device = torch.device(device)
        desc = np.zeros((len(shapes), 3), np.int64)
        off = 0
        for i, (h, w) in enumerate(shapes):
            desc[i] = (off, h, w)
            off += (h * w * 3 + 15) & ~15
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.data = torch.randint(0, 256, (off,), device=self.device, dtype=torch.uint8, generator=g)
        self.desc_host = desc
        self.desc = torch.from_numpy(desc).to(self.device)
        self.labels = np.asarray(labels)
        self.cl_list = np.unique(self.labels).tolist()
        self.sub_meta = [np.nonzero(self.labels == cl)[0] for cl in self.cl_list]
        self.nbytes = off
        return self

    @classmethod
    def from_json(cls, data_file, device):
        """The reference's file-list format (``{"image_names": [...], "image_labels": [...]}``, dataset.py:30-31);
        images are decoded once with PIL exactly as the reference opens them (dataset.py:66-67)."""
        from PIL import Image
        with open(data_file, "r") as f:
            meta = json.load(f)
        images = [np.array(Image.open(p).convert("RGB")) for p in meta["image_names"]]
        return cls(images, meta["image_labels"], device)


def draw_crops(heights, widths, gen, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """Vectorised torchvision ``RandomResizedCrop.get_params``: ten (area, aspect) proposals per image, the first that
    fits wins, else the ratio-clamped centre crop.  Returns int arrays top, left, h, w."""
    H = np.asarray(heights, np.int64)
    W = np.asarray(widths, np.int64)
    n = H.shape[0]
    u = torch.rand(n, 10, 4, generator=gen, dtype=torch.float64).numpy()
    area = (H * W).astype(np.float64)[:, None] * (scale[0] + (scale[1] - scale[0]) * u[..., 0])
    lr0, lr1 = math.log(ratio[0]), math.log(ratio[1])
    aspect = np.exp(lr0 + (lr1 - lr0) * u[..., 1])
    w = np.rint(np.sqrt(area * aspect)).astype(np.int64)
    h = np.rint(np.sqrt(area / aspect)).astype(np.int64)
    valid = (w > 0) & (w <= W[:, None]) & (h > 0) & (h <= H[:, None])
    first = np.argmax(valid, axis=1)
    has = valid.any(axis=1)
    rows = np.arange(n)
    cw, ch = w[rows, first], h[rows, first]
    top = np.floor(u[rows, first, 2] * (H - ch + 1)).astype(np.int64)
    left = np.floor(u[rows, first, 3] * (W - cw + 1)).astype(np.int64)
    # fallback: whole image clamped to the ratio range, centred
    in_ratio = W / H
    fw = np.where(in_ratio < min(ratio), W, np.where(in_ratio > max(ratio), np.rint(H * max(ratio)).astype(np.int64), W))
    fh = np.where(in_ratio < min(ratio), np.rint(W / min(ratio)).astype(np.int64), H)
    cw = np.where(has, cw, fw)
    ch = np.where(has, ch, fh)
    top = np.where(has, np.minimum(top, H - ch), (H - ch) // 2)
    left = np.where(has, np.minimum(left, W - cw), (W - cw) // 2)
    return top, left, ch, cw


def pillow_ksize(in_size, out_size):
    """Taps Pillow allocates per output pixel for the bilinear filter (Resample.c precompute_coeffs)."""
    scale = max(float(in_size) / float(out_size), 1.0)
    return int(math.ceil(scale)) * 2 + 1


class EpisodeFeeder:
    """Iterating yields ``(x, y)`` like the reference's loader -- ``x [n_way, S+Q, 3, H, W]`` fp32, ``y [n_way, S+Q]``
    class labels -- but already on the device.  ``device_packs(E)`` yields ``[E, n_way, S+Q, 3, H, W]`` tensors, one
    kernel launch each (what ``DKT.train_loop`` consumes when ``episodes_per_step = E``)."""

    SMEM_BUDGET = 100 * 1024          # two CTAs per SM

    def __init__(self, store, image_size, n_way, n_support, n_query, n_episode=100, aug=False, seed=None, lib=None):
        self.store = store
        self.S = int(image_size)
        self.n_way, self.batch = int(n_way), int(n_support) + int(n_query)
        self.n_episode = int(n_episode)
        self.aug = bool(aug)
        self.gen = torch.Generator()
        if seed is None:
            self.gen.seed()
        else:
            self.gen.manual_seed(int(seed))
        self.lib = lib
        if len(store.cl_list) < self.n_way:
            raise ValueError("the store holds %d classes, fewer than n_way = %d" % (len(store.cl_list), self.n_way))
        small = [cl for cl, idx in zip(store.cl_list, store.sub_meta) if len(idx) < self.batch]
        if small:
            raise ValueError("classes %s hold fewer than n_support + n_query = %d images" % (small[:5], self.batch))
        self.last_params = None
        self._out = None
        self._err = None

    def __len__(self):
        return self.n_episode

    # ------------------------------------------------------------------ host side: what to build
    def draw(self, E):
        """Episode composition + augmentation parameters for E episodes -> dict of numpy arrays
        (ids [E,C,SQ], labels [E,C,SQ], params [E*C*SQ, 8] int32, factors [E*C*SQ, 3] float32)."""
        st, C, SQ = self.store, self.n_way, self.batch
        ids = np.empty((E, C, SQ), np.int64)
        labels = np.empty((E, C, SQ), np.int64)
        for e in range(E):
            classes = torch.randperm(len(st.cl_list), generator=self.gen)[:C].tolist()       # dataset.py:86
            for j, ci in enumerate(classes):
                members = st.sub_meta[ci]
                pick = torch.randperm(len(members), generator=self.gen)[:SQ].numpy()           # dataset.py:47-56
                ids[e, j] = members[pick]
                labels[e, j] = st.cl_list[ci]
        flat = ids.reshape(-1)
        n = flat.shape[0]
        H, W = st.desc_host[flat, 1], st.desc_host[flat, 2]
        params = np.zeros((n, 8), np.int32)
        factors = np.ones((n, 3), np.float32)
        params[:, 0] = flat
        if self.aug:
            top, left, ch, cw = draw_crops(H, W, self.gen)
            u = torch.rand(n, 4, generator=self.gen)
            jit = torch.tensor(JITTER, dtype=torch.float32)
            factors = (jit * (u[:, :3] * 2.0 - 1.0) + 1).numpy()                              # additional_transforms.py:30
            params[:, 5] = (u[:, 3] < 0.5).numpy()
            params[:, 6] = 1
        else:
            top, left, ch, cw = np.zeros_like(H), np.zeros_like(W), H, W
        params[:, 1], params[:, 2], params[:, 3], params[:, 4] = top, left, ch, cw
        return {"ids": ids, "labels": labels, "params": params, "factors": np.ascontiguousarray(factors, np.float32)}

    def geometry(self):
        """(RH, RW, oy, ox): the crop box is resized to RH x RW and the S x S window at (oy, ox) is kept."""
        S = self.S
        if self.aug:
            return S, S, 0, 0
        big = int(S * 1.15)                                        # datamgr.py:31
        o = int(round((big - S) / 2.0))                            # torchvision CenterCrop
        return big, big, o, o

    # ------------------------------------------------------------------ device side: build it
    def transform(self, drawn, out=None, events=None):
        """Run the transform kernel for a ``draw()`` result -> fp32 [n, 3, S, S] on the store's device."""
        lib = self.lib or _lib.load()
        st, S = self.store, self.S
        dev = st.device
        RH, RW, oy, ox = self.geometry()
        p = drawn["params"]
        n = p.shape[0]
        kmax = max(pillow_ksize(int(p[:, 4].max()), RW), pillow_ksize(int(p[:, 3].max()), RH))
        kmax = (kmax + 3) & ~3
        fixed = lib.episode_transform_smem(S, kmax, 1) - S * 3
        rows_needed = int(p[:, 3].max())
        tmp_rows = max(min(rows_needed, (self.SMEM_BUDGET - fixed) // (S * 3)), min(rows_needed, 2 * kmax))
        if fixed + tmp_rows * S * 3 > 227 * 1024:
            raise _lib.DktbError("episode transform: image_size %d with %d taps does not fit shared memory" % (S, kmax))
        pin = dev.type == "cuda"
        params = torch.from_numpy(p)
        factors = torch.from_numpy(drawn["factors"])
        if pin:
            params, factors = params.pin_memory(), factors.pin_memory()
        params = params.to(dev, non_blocking=True)
        factors = factors.to(dev, non_blocking=True)
        if out is None:
            out = torch.empty(n, 3, S, S, device=dev, dtype=torch.float32)
        if self._err is None:
            self._err = torch.zeros(1, device=dev, dtype=torch.int32)
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
        if events is not None:
            events[0].record()
        lib.episode_transform(st.data, st.desc, len(st), params, factors, out, n, S, RH, RW, oy, ox, kmax, tmp_rows,
                              MEAN[0], MEAN[1], MEAN[2], STD[0], STD[1], STD[2], self._err, stream)
        if events is not None:
            events[1].record()
        self.last_params = (params, factors)          # keep the parameter buffers alive until the kernel has run
        return out

    def check(self):
        """Host-synchronising check of the kernel's error flag."""
        if self._err is not None:
            code = int(self._err.item())
            if code:
                raise _lib.DktbError("episode transform failed with code %d (1: taps > kmax, 2: band does not fit, "
                                     "3: bad image id / crop box)" % code)

    def device_packs(self, E):
        C, SQ, S = self.n_way, self.batch, self.S
        left = self.n_episode
        while left > 0:
            e = min(E, left)
            left -= e
            if self._out is None or self._out.shape[0] != e * C * SQ:
                self._out = torch.empty(e * C * SQ, 3, S, S, device=self.store.device, dtype=torch.float32)
            x = self.transform(self.draw(e), out=self._out)
            yield x.view(e, C, SQ, 3, S, S)
        self.check()

    def __iter__(self):
        C, SQ, S = self.n_way, self.batch, self.S
        for _ in range(self.n_episode):
            d = self.draw(1)
            x = self.transform(d)
            y = torch.from_numpy(d["labels"][0])
            yield x.view(C, SQ, 3, S, S), y
        self.check()


class SetDataManager:
    """``data.datamgr.SetDataManager`` (datamgr.py:67-84) on the device feeder."""

    def __init__(self, image_size, n_way, n_support, n_query, n_eposide=100, device="cuda", seed=None):
        self.image_size = image_size
        self.n_way = n_way
        self.batch_size = n_support + n_query
        self.n_support, self.n_query = n_support, n_query
        self.n_eposide = n_eposide
        self.device = device
        self.seed = seed

    def get_data_loader(self, data_file, aug):
        store = data_file if isinstance(data_file, EpisodeStore) else EpisodeStore.from_json(data_file, self.device)
        return EpisodeFeeder(store, self.image_size, self.n_way, self.n_support, self.n_query, self.n_eposide, aug=aug,
                             seed=self.seed)
