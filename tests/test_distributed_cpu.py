"""world_size-2 gloo run of the Kirchhoff output-range sharding (host logic only; the per-rank compute is
the oracle, injected as `compute`, so no GPU is needed)."""
import os
import socket
import tempfile

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


def _worker(rank, world, port, q, shm_path):
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
    # the halo exchange: every rank receives only its window of input columns; the injected column-window compute
    # asserts that it never sees more, poisons what the row chunk may not read, and the image lands on rank 0 only
    S2, T2 = 40, 400
    tt2 = np.arange(S2) * 0.01
    dk2 = np.arange(T2) * 0.0005                      # 0.5 m spacing: aperture = 68 traces << T2
    full2 = np.random.default_rng(5).standard_normal((S2, T2)).astype(np.float32)
    ref2 = om.kirchhoff(full2.astype(np.float64), tt2, dk2, 1.69e8, False)
    reach = 1.69e8 * tt2.max() * 1e-6 / 2.0

    def window_fn(snum, tt_, dk_, vel, xb, xe):
        d = np.asarray(dk_) * 1e3
        return (int(np.searchsorted(d, d[xb] - reach * 1.0001 - 1e-9, side='left')),
                int(np.searchsorted(d, d[xe - 1] + reach * 1.0001 + 1e-9, side='right')))

    wcalls = []

    def compute_window(win, c0, tnum, tt_, dk_, vel, nf, xb, xe, out_block, rows):
        r0, r1, g_hi = rows
        wcalls.append((c0, win.shape[1], r0, r1, g_hi))
        assert tnum == T2 and win.shape[1] < T2        # a window, not the image
        src_img = np.zeros((S2, T2))
        src_img[:, c0:c0 + win.shape[1]] = win.numpy().astype(np.float64)
        src_img[:max(r0 - 1, 0)] = 1e30                # rows this chunk may not read
        src_img[:, :c0] = 1e30                         # columns outside the window: must not matter
        src_img[:, c0 + win.shape[1]:] = 1e30
        out_block[r0:r1] = torch.from_numpy(om.kirchhoff(src_img, tt_, dk_, vel, nf, xb, xe)[r0:r1].astype(np.float32))

    x4 = torch.from_numpy(full2.copy()) if rank == 0 else torch.empty((1, 1)).expand(S2, T2)
    out4 = parallel.kirchhoff_sharded_device(x4, tt2, dk2, 1.69e8, False, rank=rank, world=world, gather='src',
                                             compute=compute, compute_window=compute_window, window_fn=window_fn,
                                             pipeline_chunks=3)
    if rank == 0:
        err = max(err, float(np.linalg.norm(out4.numpy() - ref2) / np.linalg.norm(ref2)))
    else:
        assert out4 is None
    assert [c[2:4] for c in wcalls] == list(reversed(parallel.row_chunks(S2, 3)))
    # the peer-mapped output image, stood in for by a file both processes map: rank 1's block is written straight
    # into "rank 0's memory", rank 0 learns from the per-chunk all_reduce that the rows are there and copies them out
    class SharedImage(object):
        def __init__(self, path):
            self.img = torch.from_file(path, shared=True, size=S2 * T2, dtype=torch.float32).view(S2, T2)
            self.copies = []

        def block(self, xb, xe):
            return self.img[:, xb:xe]

        def copy_rows_to(self, out, r0, r1, c0, c1):
            self.copies.append((r0, r1, c0, c1))
            out[r0:r1, c0:c1] = self.img[r0:r1, c0:c1]

    shared = SharedImage(shm_path)
    if rank == 0:
        shared.img.fill_(float('nan'))
    dist.barrier()
    del wcalls[:]
    out5 = parallel.kirchhoff_sharded_device(x4, tt2, dk2, 1.69e8, False, rank=rank, world=world, gather='src',
                                             compute=compute, compute_window=compute_window, window_fn=window_fn,
                                             pipeline_chunks=(1, 2, 1), peer_image=shared)
    rngs = parallel.kirchhoff_output_ranges(T2, world, tt2, dk2, 1.69e8)
    if rank == 0:
        assert np.array_equal(out5.numpy(), out4.numpy())               # same bits as the all_to_all gather
        assert shared.copies == [(r0, r1, c0, c1) for r0, r1 in reversed(parallel.row_chunks(S2, (1, 2, 1)))
                                 for c0, c1 in ((0, rngs[0][0]), (rngs[0][1], T2)) if c1 > c0]
        assert np.isnan(shared.img[:, rngs[0][0]:rngs[0][1]].numpy()).all()    # rank 0 wrote its block into `out` itself
    else:
        assert out5 is None and shared.copies == []
    assert [c[2:4] for c in wcalls] == list(reversed(parallel.row_chunks(S2, (1, 2, 1))))
    blk4, rng4 = parallel.kirchhoff_sharded_device(x4, tt2, dk2, 1.69e8, False, rank=rank, world=world, gather=False,
                                                   compute=compute, compute_window=compute_window, window_fn=window_fn,
                                                   pipeline_chunks=1)
    err = max(err, float(np.linalg.norm(blk4.numpy() - ref2[:, rng4[0]:rng4[1]]) / np.linalg.norm(ref2)))
    block, rng_ = parallel.kirchhoff_sharded_device(x, tt, dk, 1.69e8, False, rank=rank, world=world,
                                                    compute=compute, gather=False, exchange='broadcast')
    q.put((rank, err, tuple(block.shape), rng_))
    dist.destroy_process_group()


def test_kirchhoff_sharded_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shm = tempfile.NamedTemporaryFile(suffix=".img")
    shm.truncate(40 * 400 * 4)
    shm.flush()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, shm.name)) for r in range(world)]
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
