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
        # ranks that submit different job lists are caught before the gather (no hang, no misassembled result)
        try:
            run_jobs_sharded(_Eng(), _Prob(), None, jobs[:5] if rank else jobs)
            ok2 = False
        except RuntimeError as exc:
            ok2 = ok2 and "different job lists" in str(exc)
        # a whole workflow inside the process group: the G2 sweep shards over the ranks (oracle backend)
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        from oracle_backend import oracle_backend
        from pyaceqd_b200.pulses import ChirpedPulse
        from pyaceqd_b200.two_level_system.tls import tls
        from pyaceqd_b200.two_time.correlations import three_op_two_time
        p = ChirpedPulse(tau_0=0.5, e_start=0, alpha=0, t0=1.5, e0=2.0)
        with oracle_backend() as eng:
            eng.device = 0
            # sharding is opt-in: without it every rank computes the whole sweep and meets no collective
            three_op_two_time(tls, np.round(np.arange(0.0, 3.0, 0.5), 6), p, tau_max=2.0, dt=0.25,
                              options={"lindblad": True, "phonons": False, "gamma_e": 0.2})
            ok2 = ok2 and sum(c[0] for c in eng.calls) == 6
            eng.calls.clear()
            os.environ["ACEQD_DISTRIBUTED"] = "1"
            t1, tau, G = three_op_two_time(tls, np.round(np.arange(0.0, 3.0, 0.5), 6), p, tau_max=2.0, dt=0.25,
                                           options={"lindblad": True, "phonons": False, "gamma_e": 0.2})
            shard_sizes = [c[0] for c in eng.calls]
        q.put((rank, bool(ok), bool(ok2), G.shape, float(np.abs(G).sum()), shard_sizes))
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
    assert [g[:3] for g in got] == [(0, True, True), (1, True, True)]
    # both ranks hold the full, identical G2 map; each computed only its share of the 6 trajectories
    assert got[0][3] == got[1][3] == (6, 9) and abs(got[0][4] - got[1][4]) < 1e-15 and got[0][4] > 0
    assert sum(got[0][5]) + sum(got[1][5]) == 6 and all(0 < sum(g[5]) < 6 for g in got)
    from oracle_backend import oracle_backend
    from pyaceqd_b200.pulses import ChirpedPulse
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.correlations import three_op_two_time
    with oracle_backend():
        _, _, G = three_op_two_time(tls, np.round(np.arange(0.0, 3.0, 0.5), 6),
                                    ChirpedPulse(tau_0=0.5, e_start=0, alpha=0, t0=1.5, e0=2.0), tau_max=2.0, dt=0.25,
                                    options={"lindblad": True, "phonons": False, "gamma_e": 0.2})
    assert abs(float(np.abs(G).sum()) - got[0][4]) < 1e-12
