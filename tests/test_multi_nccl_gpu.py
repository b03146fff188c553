"""N>1 on real GPUs (NCCL, world_size 2, skipped on a single-GPU box): episodes sharded across ranks, one all-reduce of
the flat gradient buffer over NVLink, replicas bit-identical afterwards, and the all-reduced gradient equals the
single-process packed step over the same episodes (5-way 1-shot Conv4 at 84x84)."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model(dev, E):
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    torch.manual_seed(0)
    m = DKT(backbone.Conv4, 5, 1, kernel="bncossim", episodes_per_step=E).to(dev)
    m.monitor = False
    m.train()
    m._ensure_packed()
    m._new_adam()
    return m


def _episodes():
    from oracle import episode as oep
    return torch.stack([oep.synthetic_episode(i, 5, 1, 3, 84) for i in range(4)])


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    m = _make_model(dev, 2)
    xs = _episodes()
    out = m.train_step(xs[2 * rank:2 * rank + 2].to(dev))
    torch.cuda.synchronize()
    torch.save({"grad": m._pack.grad.cpu(), "flat": m._pack.flat.cpu(), "bufs": m._bufs.flat.cpu(),
                "loss": out["loss"].cpu()}, os.path.join(outdir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_match_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, "rank0.pt"))
        r1 = torch.load(os.path.join(d, "rank1.pt"))
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["grad"], r1["grad"])
    assert torch.equal(r0["bufs"], r1["bufs"])
    dev = torch.device("cuda", 0)
    m = _make_model(dev, 4)
    out = m.train_step(_episodes().to(dev))
    g_single = m._pack.grad.cpu()
    g_dist = r0["grad"] / 2.0
    err = float((g_single - g_dist).abs().max() / g_single.abs().max())
    assert err < 1e-5, err
    losses = torch.cat([r0["loss"], r1["loss"]])
    assert torch.allclose(losses, out["loss"].cpu(), rtol=1e-6, atol=1e-7)
    mask = g_single.abs() > 1e-3 * g_single.abs().max()
    assert float((m._pack.flat.cpu() - r0["flat"])[mask].abs().max()) < 1e-6
