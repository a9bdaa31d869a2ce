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


def _adam_worker(rank, world, port, results):
    """dmgs_adam_exchange_peer against all-reduce + FusedAdam on every replica (three steps, both transports)."""
    import torch.distributed as dist
    from dmgs_b200 import multiview as MV
    from dmgs_b200.optim import FusedAdam
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        P = 50_001  # odd: fields with a padded tail
        widths = {"means3D": (3,), "opacities": (1,), "scales": (3,), "rotations": (4,), "shs": (16, 3)}
        groups = {"means3D": {"lr": 1.6e-4}, "opacities": {"lr": 5e-2}, "scales": {"lr": 5e-3}, "rotations": {"lr": 1e-3},
                  "shs": {"lr": 2.5e-3, "lr_hi": 2.5e-3 / 20, "period": 48, "split": 3}}
        out = {}
        transports = [False, True]
        for use_mc in transports:
            galloc, gh = MV.SymmetricFlat.allocator(dev, dist.group.WORLD)
            grads = MV.FlatGradBuffer(P, MV.RASTER_WIDTHS_SH, dev, allocate=galloc)
            palloc, ph = MV.SymmetricFlat.allocator(dev, dist.group.WORLD)
            params = MV.FlatGradBuffer(P, widths, dev, allocate=palloc)
            if use_mc and not (gh[0].multicast_ptr and ph[0].multicast_ptr):
                continue
            gen = torch.Generator(device=dev).manual_seed(7)  # the same parameters on every rank
            for name in widths:
                params.views[name].copy_(torch.randn(params.views[name].shape, generator=gen, device=dev))
            ref_p = {k: params.views[k].clone() for k in widths}
            ref_opt = FusedAdam([{"params": [ref_p[k]], "name": k, **groups[k]} for k in widths], lr=0.0, eps=1e-15)
            opt = MV.ShardedPeerAdam(params, ph[0], grads, gh[0], groups, eps=1e-15)
            worst = 0.0
            for it in range(3):
                g = torch.Generator(device=dev).manual_seed(1000 * it + rank)  # different gradients on every rank
                grads.flat.zero_()  # the padding between the fields is never written by the backward: it stays zero
                for name in grads.views:
                    grads.views[name].copy_(torch.randn(grads.views[name].shape, generator=g, device=dev))
                ref_g = grads.flat.clone()
                dist.all_reduce(ref_g)
                ref_views = {name: ref_g[o:o + m].view(P, *MV.RASTER_WIDTHS_SH[name]) for name, (o, m) in grads.offsets.items()}
                ref_opt.step(grads={k: ref_views[k] for k in widths}, grad_scale=0.25)
                opt.step(grad_scale=0.25, use_multicast=use_mc)
                torch.cuda.synchronize()
                for k in widths:
                    worst = max(worst, float((params.views[k] - ref_p[k]).abs().max() / ref_p[k].abs().max()))
            key = "multimem" if use_mc else "p2p"
            out[key] = worst
            mine = params.flat.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            out[key + "_same_bits"] = bool(torch.equal(mine, other))
            # the padding floats between the fields stay zero
            pad = torch.ones(params.flat.numel(), dtype=torch.bool, device=dev)
            for o, m in params.offsets.values():
                pad[o:o + m] = False
            out[key + "_pad_zero"] = bool((params.flat[pad] == 0).all())
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_adam_exchange_matches_allreduce_then_adam():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_adam_worker, args=(world, 29534, results), nprocs=world, join=True)
        res = dict(results)
    assert len(res) == world
    for rank, out in res.items():
        assert "p2p" in out
        for k, v in out.items():
            if k.endswith("_same_bits") or k.endswith("_pad_zero"):
                assert v, (rank, k)
            else:
                assert v <= 1e-6, (rank, k, v)
