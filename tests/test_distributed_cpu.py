"""world_size-2 gloo run of the Kirchhoff output-range sharding (host logic only; the per-rank compute is
the oracle, injected as `compute`, so no GPU is needed)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from impdar_b200 import parallel
    from oracle import migration as om
    S, T = 24, 40
    tt = np.arange(S) * 0.01
    dk = np.arange(T) * 0.005
    rng = np.random.default_rng(3)
    full = rng.standard_normal((S, T)).astype(np.float32)
    x = torch.from_numpy(full.copy()) if rank == 0 else torch.zeros((S, T), dtype=torch.float32)

    def compute(xt, tt_, dk_, vel, nf, xb, xe):
        return torch.from_numpy(om.kirchhoff(xt.numpy().astype(np.float64), tt_, dk_, vel, nf, xb, xe).astype(np.float32))

    out = parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world, compute=compute)
    ref = om.kirchhoff(full.astype(np.float64), tt, dk, 1.69e8, False)
    err = float(np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref))

    # the pipelined exchange: bottom-up row chunks.  The injected row-range compute poisons every input row the
    # chunk is not allowed to read (rows < s_begin - 1), so a wrong dependency claim would wreck the result.
    calls = []

    def compute_rows(xt, tt_, dk_, vel, nf, xb, xe, r0, r1, g_hi, out_block):
        calls.append((r0, r1, g_hi))
        src = xt.numpy().astype(np.float64).copy()
        src[:max(r0 - 1, 0)] = 1e30
        out_block[r0:r1] = torch.from_numpy(om.kirchhoff(src, tt_, dk_, vel, nf, xb, xe)[r0:r1].astype(np.float32))

    x2 = torch.from_numpy(full.copy()) if rank == 0 else torch.zeros((S, T), dtype=torch.float32)
    out2 = parallel.kirchhoff_sharded_device(x2, tt, dk, 1.69e8, False, rank=rank, world=world, compute=compute,
                                             compute_rows=compute_rows, pipeline_chunks=5)
    err = max(err, float(np.linalg.norm(out2.numpy() - ref) / np.linalg.norm(ref)))
    assert [c[:2] for c in calls] == list(reversed(parallel.row_chunks(S, 5)))
    assert [c[2] for c in calls] == [S] + [c[0] for c in calls[:-1]]          # g_hi chains bottom-up
    assert np.array_equal(x2.numpy(), full)                                  # every rank ends up with the whole input

    def irregular(*a):
        raise ValueError('kirchhoff_rows: row-range calls need uniform trace spacing (the table path)')

    x3 = torch.from_numpy(full.copy()) if rank == 0 else torch.zeros((S, T), dtype=torch.float32)
    out3 = parallel.kirchhoff_sharded_device(x3, tt, dk, 1.69e8, False, rank=rank, world=world, compute=compute,
                                             compute_rows=irregular, pipeline_chunks=4)
    err = max(err, float(np.linalg.norm(out3.numpy() - ref) / np.linalg.norm(ref)))
    block, rng_ = parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world,
                                                    compute=compute, gather=False)
    q.put((rank, err, tuple(block.shape), rng_))
    dist.destroy_process_group()


def test_kirchhoff_sharded_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(e < 1e-6 for _, e, _, _ in res)
    (b0, e0), (b1, e1) = res[0][3], res[1][3]
    assert b0 == 0 and e0 == b1 and e1 == 40
    assert res[0][2] == (24, e0 - b0) and res[1][2] == (24, e1 - b1)
