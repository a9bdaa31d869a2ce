"""Host logic of the view-partitioned step (dmgs_b200/multiview.py) on CPU: world_size-2 gloo ranks,
the oracle playing the renderer.  The N-rank result must equal the 1-rank result on the whole batch
(to fp32 summation-order tolerance)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmgs_b200 import multiview as MV
from dmgs_b200 import synthetic as S


def test_partition_covers_every_view_once():
    for n_views in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            seen = sorted(v for r in range(world) for v in MV.partition_views(n_views, world, r))
            assert seen == list(range(n_views))
            sizes = [len(MV.partition_views(n_views, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        MV.partition_views(4, 2, 2)


def test_flat_buffer_layout():
    buf = MV.FlatGradBuffer(10, MV.RASTER_WIDTHS_SH, "cpu")
    assert buf.views["shs"].shape == (10, 16, 3) and buf.views["rotations"].shape == (10, 4)
    assert buf.flat.numel() >= 10 * (3 + 3 + 1 + 3 + 4 + 48)
    for name, (o, n) in buf.offsets.items():
        assert o % 4 == 0, f"{name} is not 16-byte aligned"
    buf.views["opacities"].fill_(2.0)
    o, n = buf.offsets["opacities"]
    assert torch.all(buf.flat[o:o + n] == 2.0) and buf.flat.sum() == 2.0 * 10
    assert buf.zero_().flat.abs().sum() == 0


def _oracle_view(v, acc, cl_np, W, H):
    """Oracle forward+backward of view v, gradients added into acc (numpy-backed CPU tensors)."""
    from oracle import oracle as O
    cam = S.nerf_synthetic_camera(v, W, H)
    P = cl_np["means3D"].shape[0]
    pr = O.make_params(P, W, H, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), [0, 0, 0],
                       cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.camera_center.numpy())
    kw = dict(scales=cl_np["scales"], rotations=cl_np["rotations"], shs=cl_np["shs"])
    fwd = O.render_forward(pr, cl_np["means3D"], cl_np["opacities"], **kw)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + v)).numpy()
    bw = O.render_backward(pr, fwd, dL, cl_np["means3D"], **kw)
    acc["means3D"] += torch.from_numpy(bw["dL_dmeans3D"])
    acc["means2D"][:, :2] += torch.from_numpy(bw["dL_dmean2D"])
    acc["opacities"][:, 0] += torch.from_numpy(bw["dL_dopacity"])
    acc["scales"] += torch.from_numpy(bw["dL_dscales"])
    acc["rotations"] += torch.from_numpy(bw["dL_drotations"])
    acc["shs"] += torch.from_numpy(bw["dL_dshs"])
    return torch.tensor(float((fwd["img"]["color"] * dL).sum()))


def _scene():
    cl = S.random_cloud(400, seed=5, extent=1.0, log_scale_mean=math.log(0.06))
    return {k: v.numpy() for k, v in cl.items()}


N_VIEWS, W, H = 5, 96, 64


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cl = _scene()
        buf = MV.FlatGradBuffer(400, MV.RASTER_WIDTHS_SH, "cpu")
        vp = MV.ViewParallel()
        assert vp.world == world and vp.rank == rank
        loss = vp.step(N_VIEWS, lambda v, acc: _oracle_view(v, acc, cl, W, H), buf, average=True)
        np.save(os.path.join(out_dir, f"flat{rank}.npy"), buf.flat.numpy())
        np.save(os.path.join(out_dir, f"loss{rank}.npy"), loss.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cl = _scene()
    buf = MV.FlatGradBuffer(400, MV.RASTER_WIDTHS_SH, "cpu")
    loss = MV.ViewParallel().step(N_VIEWS, lambda v, acc: _oracle_view(v, acc, cl, W, H), buf, average=True)
    ref = buf.flat.numpy()
    f0, f1 = np.load(tmp_path / "flat0.npy"), np.load(tmp_path / "flat1.npy")
    assert np.array_equal(f0, f1), "ranks disagree after the all-reduce"
    assert np.abs(ref).max() > 0
    assert np.linalg.norm(f0 - ref) <= 1e-5 * np.linalg.norm(ref)
    l0, l1 = float(np.load(tmp_path / "loss0.npy")), float(np.load(tmp_path / "loss1.npy"))
    assert l0 == l1 and abs(l0 - float(loss)) <= 1e-4 * abs(float(loss))


def test_shard_rows_cover_every_row_once():
    """StagedInputs at N > 1: equal chunks (all-gather), every row staged by exactly one rank, ragged tails."""
    from dmgs_b200.multiview import shard_rows
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            seen = 0
            chunks = set()
            for r in range(world):
                lo, hi, chunk = shard_rows(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= chunk
                assert lo == min(r * chunk, n)
                seen += hi - lo
                chunks.add(chunk)
            assert seen == n and len(chunks) == 1 and chunks.pop() * world >= n
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_adam_exchange_shards_partition_every_field():
    """dmgs_adam_exchange_shard: the ranks' 16-byte-group ranges tile ceil(n/4) groups exactly once, in rank order."""
    import ctypes as C
    from dmgs_b200 import _lib as L
    lib = L.lib()
    for n in (0, 1, 5, 50_001, 150_003, 48_000_000):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                b, e = C.c_int64(), C.c_int64()
                assert lib.dmgs_adam_exchange_shard(n, world, r, C.byref(b), C.byref(e)) == 0
                assert 0 <= b.value <= e.value and b.value in (prev, e.value)
                prev = e.value
            assert prev == (n + 3) // 4
    b, e = C.c_int64(), C.c_int64()
    assert lib.dmgs_adam_exchange_shard(10, 2, 2, C.byref(b), C.byref(e)) != 0
