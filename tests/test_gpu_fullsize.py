"""BASELINE.json's full cfg2 size on the GPU (4096 pulse-area x detuning trajectories x 400 steps, chi = 128) through
the public sweep call, checked by properties that do not need the oracle to run 1.6 million trajectory-steps:
schedule and transport independence (bit for bit), a sample of trajectories against the oracle, linearity in the
initial state, and the small-bond kernel at the same batch size."""
import os
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _jobs_of(tables, idx, n_steps, dt, rho0=None):
    from pyaceqd_b200.jobs import FieldTable, Job
    return [Job(0.0, n_steps * dt, dt, tables={"x": FieldTable(0.0, dt, tables[i, 0])}, rho0=rho0) for i in idx]


@pytest.mark.parametrize("chi", [128, 16])
def test_cfg2_full_size_properties(engine, chi, monkeypatch):
    import bench
    n_steps, dt = 400, 0.1
    prob, pt, tables = bench.make_workload(chi, 64, 64, n_steps, dt)
    n_traj = tables.shape[0]
    assert n_traj == 4096
    for k in ("ACEQD_SEGMENTS", "ACEQD_NO_ZEROCOPY", "ACEQD_SMALL"):
        monkeypatch.delenv(k, raising=False)
    fast = engine.run_sweep(prob, pt, tables, (0.0, dt), 0.0, n_steps, dt)      # segments / small kernel, zero copy
    assert fast.shape == (n_traj, n_steps + 1, prob.n_out)
    assert np.isfinite(fast.view(np.float64)).all()
    # 1. same numbers from the wave schedule + staged copy (and from the tile kernel for small bonds)
    monkeypatch.setenv("ACEQD_SEGMENTS", "0")
    monkeypatch.setenv("ACEQD_NO_ZEROCOPY", "1")
    slow = engine.run_sweep(prob, pt, tables, (0.0, dt), 0.0, n_steps, dt)
    if chi > 32:
        assert np.array_equal(fast, slow)
    else:
        monkeypatch.setenv("ACEQD_SMALL", "0")
        tile = engine.run_sweep(prob, pt, tables, (0.0, dt), 0.0, n_steps, dt)
        assert np.array_equal(fast, slow)                    # transport only
        assert np.abs(fast - tile).max() < 1e-12             # other kernel: same arithmetic up to summation order
    for k in ("ACEQD_SEGMENTS", "ACEQD_NO_ZEROCOPY", "ACEQD_SMALL"):
        monkeypatch.delenv(k, raising=False)
    # 2. a sample of trajectories against the oracle (first, around the tile cut of the segment schedule, last)
    idx = [0, 1, 16 * 147 + 3, 16 * 148 + 9, 2049, 4095]
    for i, jb in zip(idx, _jobs_of(tables, idx, n_steps, dt)):
        ref = oracle.propagate(prob, pt, jb)                 # [n_out, n_steps + 1]
        assert np.abs(fast[i].T - ref).max() < TOL, i
    # 3. linearity in the initial state: rho0 = a rho1 + b rho2  ->  a out1 + b out2
    rng = np.random.default_rng(3)
    r1 = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    r2 = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    a, b = 0.3 - 0.7j, -1.1 + 0.2j
    sub = [5, 777, 4000]
    o1 = engine.run_jobs(prob, pt, _jobs_of(tables, sub, n_steps, dt, r1))
    o2 = engine.run_jobs(prob, pt, _jobs_of(tables, sub, n_steps, dt, r2))
    o3 = engine.run_jobs(prob, pt, _jobs_of(tables, sub, n_steps, dt, a * r1 + b * r2))
    for x, y, z in zip(o1, o2, o3):
        assert np.abs(a * x + b * y - z).max() < 1e-11 * max(1.0, np.abs(z).max())
