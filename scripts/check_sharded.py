"""torchrun check of the multi-GPU Kirchhoff: the pipelined exchange (bottom-up row chunks: broadcast | kernels |
all-gather overlapped) must return bit for bit what the three phases back to back return; prints both timings.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_sharded.py [S T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from impdar_b200 import parallel, synthetic

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
S, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 16384)
tt, dk, _ = synthetic.geometry(S, T)
full = synthetic.diffractor_radargram(S, T, seed=9, n_diffractors=64)

def run(chunks):
    x = full.clone() if rank == 0 else torch.zeros((S, T), dtype=torch.float32, device="cuda")
    return parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world, pipeline_chunks=chunks), x

def timed(chunks, n=3):
    run(chunks); torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(n):
        x = full.clone() if rank == 0 else torch.zeros((S, T), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world, pipeline_chunks=chunks)
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(t.item())
    return min(ts)

ref, _ = run(1)
for chunks in (4, 8, 16):
    got, x = run(chunks)
    same = bool(torch.equal(got, ref)) and bool(torch.equal(x, full))
    t = torch.tensor([0.0 if same else 1.0], device="cuda"); dist.all_reduce(t)
    if rank == 0:
        print("pipeline_chunks=%d identical on all ranks: %s" % (chunks, t.item() == 0.0), flush=True)
for chunks in (1, 4, 8, 16):
    ms = timed(chunks)
    if rank == 0:
        print("%d x %d on %d GPUs, pipeline_chunks=%d: %.2f ms" % (S, T, world, chunks, ms), flush=True)
dist.destroy_process_group()
