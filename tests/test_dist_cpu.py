"""N > 1 path on CPU: world_size-2 gloo run of the sharding + final-gather logic bench.py uses
(stars are independent: there is no data-path collective to test, only ownership and the gather)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from dstar_b200.dist import shard_range


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dstar_b200.dist import allmax, gather_monitors
    lo, hi = shard_range(n_total, rank, world)
    local = np.arange(lo, hi, dtype=np.float64)[None, :] * np.array([[1.0], [10.0], [100.0]])  # (3 epochs, n_local)
    full = gather_monitors(dist, local, n_total, rank, world)
    mx = allmax(dist, 1.5 + rank)
    if rank == 0:
        q.put((full, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (1, 7, 256, 8192):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_total, dtype=np.float64)[None, :] * np.array([[1.0], [10.0], [100.0]])
    np.testing.assert_array_equal(full, expect)
    assert mx == 2.5
