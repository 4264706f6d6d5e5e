"""N>1 path on CPU: world_size-2 gloo run of the scan sharding / result gathering / max-time logic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidiff_b200.sharding import gather_scans, max_over_ranks, scans_of_rank


def test_assignment_covers_every_scan_once():
    for n, w in ((8, 1), (8, 2), (8, 8), (5, 4), (3, 8)):
        seen = sorted(b for r in range(w) for b in scans_of_rank(n, w, r))
        assert seen == list(range(n))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_scans = 5
    local = {}
    for b in scans_of_rank(n_scans, world, rank):
        g = torch.Generator().manual_seed(b)
        local[b] = torch.randn(100 + 7 * b, 3, generator=g)          # variable lengths (postprocess filters points)
    dist.barrier()
    t = max_over_ranks(10.0 + rank, "cpu")
    res = gather_scans(local, n_scans, "cpu")
    if rank == 0:
        ok = t == 10.0 + world - 1 and sorted(res) == list(range(n_scans))
        for b in range(n_scans):
            g = torch.Generator().manual_seed(b)
            ok = ok and torch.equal(res[b], torch.randn(100 + 7 * b, 3, generator=g))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
