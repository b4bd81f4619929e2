"""BatchExecutor / run_requests host logic (CPU, oracle backend): any callable the reference's ThreadPoolExecutor
accepts must be accepted (reference fan-out idiom: two_level_system/rabi_rotations.py:172-203), requests sampled from
different start times must not abort the sweep, and multi-rank sharding is opt-in."""
import numpy as np
import pytest

from oracle_backend import oracle_backend
from pyaceqd_b200.batch import BatchExecutor, wait
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.two_level_system.tls import tls

OPTS = dict(dt=0.1, lindblad=True, phonons=False, gamma_e=0.05)


def _final_x(t0, t1, pulse, **kw):
    """An adapter that post-processes the system() result, like `t,g,x,_,_ = tls(...)` in user scripts."""
    t, g, x, pgx, pxg = tls(t0, t1, pulse, **kw)
    return float(np.real(x[-1]))


def _indexing_adapter(t0, t1, pulse, **kw):
    return tls(t0, t1, pulse, **kw)[2][-1]


def test_post_processing_adapters_fall_back_to_eager_runs():
    p = ChirpedPulse(tau_0=1.0, e_start=0.0, alpha=0, t0=3.0, e0=1.0)
    with oracle_backend() as eng:
        direct = tls(0, 6, p, **OPTS)
        with BatchExecutor(max_workers=4) as ex:
            f1 = ex.submit(_final_x, 0, 6, p, **OPTS)
            f2 = ex.submit(_indexing_adapter, 0, 6, p, **OPTS)
            f3 = ex.submit(tls, 0, 6, p, **OPTS)          # plain pass-through adapter: deferred
            wait([f1, f2, f3])
        assert abs(f1.result() - direct[2][-1].real) < 1e-14
        assert abs(f2.result() - direct[2][-1]) < 1e-14
        assert np.array_equal(f3.result(), direct)
        # a genuine error of the callable is not swallowed
        with pytest.raises(ZeroDivisionError):
            BatchExecutor().submit(lambda: 1 / 0)


def test_mixed_start_times_in_one_executor():
    """Two submits with different t_start sample their drives from different origins: the reference runs them
    independently; here they become separate batches instead of aborting the sweep."""
    p = ChirpedPulse(tau_0=1.0, e_start=0.0, alpha=0, t0=4.0, e0=1.0)
    with oracle_backend() as eng:
        a = tls(0, 6, p, **OPTS)
        b = tls(1.05, 6.05, p, **OPTS)                    # not a whole number of steps after 0
        with BatchExecutor() as ex:
            fa = ex.submit(tls, 0, 6, p, **OPTS)
            fb = ex.submit(tls, 1.05, 6.05, p, **OPTS)
            fc = ex.submit(tls, 0, 4, p, **OPTS)
        assert np.array_equal(fa.result(), a) and np.array_equal(fb.result(), b)
        assert fc.result().shape[1] == 41
        sizes = sorted(c[0] for c in eng.calls[-2:])
        assert sizes == [1, 2]                             # {0 -> 6, 0 -> 4} batched, the shifted start on its own


def test_engine_plan_rejects_off_grid_start_times():
    """Engine.plan keeps several start times in one batch only when they sit on the common step grid."""
    import sys
    from helpers import make_tables, tls_problem
    from pyaceqd_b200.engine import Engine
    from pyaceqd_b200.jobs import Job
    from pyaceqd_b200.process_tensor import trivial_pt
    prob = tls_problem(phonons=False)
    p = ChirpedPulse(tau_0=1.0, e_start=0.0, alpha=0, t0=4.0, e0=1.0)
    tabs = make_tables([p], 0.0, 8.0, 0.1)
    eng = Engine.__new__(Engine)           # planner only: no CUDA context

    class _NoGpu:
        def aceqd_max_tile(self, *a):
            return 16

        def aceqd_pass_load(self, *a):
            return 1
    eng.lib = _NoGpu()
    eng.problem_handle = lambda prob, pt: (None, np.zeros(prob.NL, dtype=np.int32))
    jobs = [Job(0.0, 2.0, 0.1, tables=tabs), Job(0.25, 2.05, 0.1, tables=tabs)]
    with pytest.raises(ValueError, match="whole number of steps"):
        eng.plan(prob, trivial_pt(len(prob.cls_keys)), jobs)
    ok = [Job(0.0, 2.0, 0.1, tables=tabs), Job(0.3, 2.0, 0.1, tables=tabs)]
    eng.plan(prob, trivial_pt(len(prob.cls_keys)), ok)
    eng.ctx = None
