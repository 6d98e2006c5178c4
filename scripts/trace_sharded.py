"""Timeline of one sharded Kirchhoff step per rank (IMPDAR_TRACE_SHARDED=1): ms since the start of the call at which each
row chunk's input window has arrived, each chunk is computed, and the image is assembled on rank 0.
    IMPDAR_TRACE_SHARDED=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/trace_sharded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from impdar_b200 import parallel, synthetic
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
S, T = 8192, 65536
tt, dk, _ = synthetic.geometry(S, T)
x = synthetic.diffractor_radargram(S, T, seed=5, n_diffractors=64) if rank == 0 else torch.empty((1, 1), device="cuda").expand(S, T)
os.environ.pop("IMPDAR_TRACE_SHARDED", None)
for _ in range(3):
    parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world, gather='src')
torch.cuda.synchronize(); dist.barrier()
os.environ["IMPDAR_TRACE_SHARDED"] = "1"
for chunks in (parallel.DEFAULT_CHUNKS, 4, 1):
    if rank == 0:
        print("pipeline_chunks =", chunks, flush=True)
    for _ in range(2):
        torch.cuda.synchronize(); dist.barrier()
        parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world, gather='src', pipeline_chunks=chunks)
dist.destroy_process_group()
