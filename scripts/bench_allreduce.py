"""All-reduce of the 248 MB flat gradient buffer (1 M Gaussians x 62 floats): NCCL vs the peer-memory kernel
(P2P loads/stores, NVSwitch multimem).  torchrun --nproc-per-node N scripts/bench_allreduce.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dmgs_b200 import multiview as MV

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
vs = MV.ViewStreams(1_000_000, MV.RASTER_WIDTHS_SH, dev, n=1, peer_group=dist.group.WORLD)
flat = vs.buf.flat
other = torch.zeros_like(flat)

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

if "--sweep" in sys.argv:  # launch shapes of the multimem kernel (DMGS_AR_SHAPE: CTAs per SM, threads, groups in flight)
    for shape in ["8,256,4", "2,512,8", "1,1024,8", "4,512,4", "2,1024,4", "16,256,2", "1,512,8", "4,256,8", "1,256,8", "2,256,8"]:
        os.environ["DMGS_AR_SHAPE"] = shape
        ms = timed(lambda: vs.peer.all_reduce_(1.0, use_multicast=True), iters=10)
        if rank == 0:
            print(json.dumps({"world": world, "shape": shape, "multimem_ms": round(ms, 4),
                              "algbw_GBps": round(flat.numel() * 4 / ms / 1e6, 1)}), flush=True)
    dist.destroy_process_group()
    sys.exit(0)
res = {"world": world, "bytes": flat.numel() * 4, "peer_error": vs.peer_error}
res["nccl_ms"] = timed(lambda: dist.all_reduce(other))
if vs.peer is not None:
    res["p2p_ms"] = timed(lambda: vs.peer.all_reduce_(1.0, use_multicast=False))
    if vs.peer.multicast_ptr:
        res["multimem_ms"] = timed(lambda: vs.peer.all_reduce_(1.0, use_multicast=True))
    res["barrier_pair_ms"] = timed(lambda: (vs.peer.hdl.barrier(channel=0), vs.peer.hdl.barrier(channel=1)))
if rank == 0:
    for k in ("nccl_ms", "p2p_ms", "multimem_ms"):
        if k in res:
            res[k.replace("_ms", "_algbw_GBps")] = round(res["bytes"] / res[k] / 1e6, 1)
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
