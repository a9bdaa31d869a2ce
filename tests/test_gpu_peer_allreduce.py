"""Peer-memory all-reduce (dmgs_allreduce_peer) against NCCL on the same buffers.  Needs >= 2 GPUs on the
box (skipped otherwise): spawns one process per GPU, like torchrun."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    import torch.distributed as dist
    from dmgs_b200 import multiview as MV
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        P = 100_003
        vs = MV.ViewStreams(P, MV.RASTER_WIDTHS_SH, dev, n=2, peer_group=dist.group.WORLD)
        assert vs.peer is not None, f"symmetric memory unavailable: {vs.peer_error}"
        n = vs.buf.flat.numel()
        out = {}
        for use_mc in ([True, False] if vs.peer.multicast_ptr else [False]):
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            vs.buf.flat.copy_(torch.randn(n, generator=g, device=dev))
            ref = vs.buf.flat.clone()
            dist.all_reduce(ref)
            ref *= 0.125
            vs.peer.all_reduce_(scale=0.125, use_multicast=use_mc)
            torch.cuda.synchronize()
            err = float((vs.buf.flat - ref).abs().max() / ref.abs().max())
            out["multimem" if use_mc else "p2p"] = err
            # every rank holds the same bits afterwards
            mine = vs.buf.flat.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            out[("multimem" if use_mc else "p2p") + "_same_bits"] = bool(torch.equal(mine, other))
        # twice in a row on the same buffer (barrier channels are reusable)
        vs.buf.flat.fill_(float(rank + 1))
        vs.all_reduce_()
        vs.all_reduce_()
        torch.cuda.synchronize()
        s = world * (world + 1) / 2
        out["twice"] = bool((vs.buf.flat == s * world).all())
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_peer_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, 29533, results), nprocs=world, join=True)
        res = dict(results)
    assert len(res) == world
    for rank, out in res.items():
        for k, v in out.items():
            if k.endswith("_same_bits") or k == "twice":
                assert v, (rank, k)
            else:
                assert v <= 1e-6, (rank, k, v)
