"""N>1 host logic on the CPU (gloo, world_size 2): sharding + the single final gather."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from pyaceqd_b200.distributed import all_gather_blocks, balanced_blocks


def test_balanced_blocks_cover_and_balance():
    costs = [10, 1, 1, 1, 9, 2, 8, 3, 7, 4]
    for world in (1, 2, 3, 4, 8, 16):
        blocks = balanced_blocks(costs, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == len(costs)
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    b2 = balanced_blocks(costs, 2)
    s = [sum(costs[a:b]) for a, b in b2]
    assert abs(s[0] - s[1]) <= max(costs)
    assert balanced_blocks([], 2) == [(0, 0), (0, 0)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ragged complex blocks: rank r holds counts[r] rows of shape [3]
        counts = [5, 2]
        rng = np.random.default_rng(100 + rank)
        local = rng.standard_normal((counts[rank], 3)) + 1j * rng.standard_normal((counts[rank], 3))
        full = all_gather_blocks(local, counts)
        exp = []
        for r in range(world):
            g = np.random.default_rng(100 + r)
            exp.append(g.standard_normal((counts[r], 3)) + 1j * g.standard_normal((counts[r], 3)))
        ok = np.array_equal(full, np.concatenate(exp, axis=0))

        # run_jobs_sharded with a stand-in engine (host logic only; the CUDA engine needs a GPU)
        from pyaceqd_b200.distributed import run_jobs_sharded

        class _Job:
            def __init__(self, n):
                self.n_steps = n

        class _Prob:
            n_out = 2

        class _Eng:
            def run_jobs(self, prob, pt, jobs, **kw):
                return [np.full((2, j.n_steps + 1), j.n_steps + 0.5j) for j in jobs]

        jobs = [_Job(n) for n in (3, 9, 1, 4, 4, 7)]
        res = run_jobs_sharded(_Eng(), _Prob(), None, jobs)
        ok2 = all(r.shape == (2, j.n_steps + 1) and np.all(r == j.n_steps + 0.5j) for r, j in zip(res, jobs))
        q.put((rank, bool(ok), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(30)
    assert got == [(0, True, True), (1, True, True)]
