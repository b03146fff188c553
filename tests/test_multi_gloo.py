"""N>1 host logic on CPU (gloo, world_size 2): episodes sharded across ranks, ONE all-reduce of the flat gradient
buffer, identical replicas afterwards, and the result equals the single-process step over the same episodes.
Runs the kernels through the g++ emulation library (test infrastructure) -- the NCCL path on GPUs uses the
same DKT.train_step code with backend "nccl"."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model(lib, E):
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    torch.manual_seed(0)
    m = DKT(lambda: backbone.ConvNet(4, image_size=16), 2, 1, kernel="bncossim", episodes_per_step=E, lib=lib)
    m.monitor = False
    m.train()
    m._ensure_packed()
    m._new_adam()
    return m


def _episodes():
    from oracle import episode as oep
    return torch.stack([oep.synthetic_episode(i, 2, 1, 1, 16) for i in range(2)])


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import torch.distributed as dist
    import build_emu
    from deep_kernel_transfer_b200._lib import DktbLib
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = DktbLib(build_emu.build())
    m = _make_model(lib, 1)
    xs = _episodes()
    out = m.train_step(xs[rank:rank + 1])
    torch.save({"grad": m._pack.grad.clone(), "flat": m._pack.flat.clone(), "bufs": m._bufs.flat.clone(),
                "loss": out["loss"].clone()}, os.path.join(outdir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_ranks_match_single_process():
    import build_emu
    from deep_kernel_transfer_b200._lib import DktbLib
    lib = DktbLib(build_emu.build())
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "rank0.pt"))
        r1 = torch.load(os.path.join(d, "rank1.pt"))
    # replicas identical after the step
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["grad"], r1["grad"])
    assert torch.equal(r0["bufs"], r1["bufs"])
    # single process, both episodes packed: gradient of the mean loss == all-reduced gradient / world
    m = _make_model(lib, 2)
    out = m.train_step(_episodes())
    g_single = m._pack.grad
    g_dist = r0["grad"] / 2.0
    err = float((g_single - g_dist).abs().max() / g_single.abs().max())
    assert err < 1e-5, err
    losses = torch.cat([r0["loss"], r1["loss"]])
    assert torch.allclose(losses, out["loss"], rtol=1e-6, atol=1e-7)
    # parameters after Adam agree wherever the gradient sign is numerically determined
    mask = g_single.abs() > 1e-3 * g_single.abs().max()
    assert float((m._pack.flat - r0["flat"])[mask].abs().max()) < 1e-6
