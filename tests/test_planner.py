"""The array planner (pyaceqd_b200/planner.py): multi-level forking checked on the CPU by executing the planned
levels with the oracle's arithmetic (tests/plan_interpreter.py) against un-forked oracle runs of every job."""
import copy
import time

import numpy as np
import pytest

import oracle
from helpers import biexciton_problem, make_tables, tls_problem
from plan_interpreter import run_plan
from pyaceqd_b200 import planner
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.process_tensor import synthetic_pt, trivial_pt
from pyaceqd_b200.pulses import ChirpedPulse


def _check(prob, pt, jobs, fork=True, tol=1e-11):
    plan = planner.plan_levels(planner.arrays_from_jobs(prob, jobs), prob.n_out, fork=fork)
    got = run_plan(prob, pt, jobs, plan)
    for jb, g in zip(jobs, got):
        pj = prob
        if jb.rho0 is not None:
            pj = copy.copy(prob)
            pj.rho0 = np.asarray(jb.rho0, dtype=complex).reshape(-1)
        want = oracle.propagate(pj, pt, jb)
        want = want[:, -jb.tail_rows:] if jb.tail_rows else want
        assert g.shape == want.shape and np.abs(g - want).max() < tol
    return plan


def _mto(prob, op, side, t, before=False):
    return prob.parse_mtos([{"operator": op, "applyFrom": side, "applyBefore": "true" if before else "false", "time": t}])


def test_triangular_sweep_forks_twice():
    """Operators at t1, t2 and t1 + tb (reference timebin/twophoton_new.py:515-557): the stretch t1 -> t2 runs once
    per t1, the jobs start at t2."""
    prob = tls_problem()
    pt = synthetic_pt(6, len(prob.cls_keys), kind="unitary", scale=0.999)
    p = ChirpedPulse(tau_0=0.3, e_start=0.5, alpha=0, t0=0.8, e0=3.0)
    tabs = make_tables([p], 0.0, 4.0, 0.1)
    t1 = [0.2, 0.5, 0.9, 1.4]
    tb = 1.5
    jobs = []
    for i, a in enumerate(t1):
        for b in t1[i:]:
            mt = _mto(prob, "|0><1|_2", "_left", a) + _mto(prob, "|1><0|_2", "_right", b) + \
                _mto(prob, "|0><1|_2", "_right", a + tb)
            jobs.append(Job(0.0, b + tb, 0.1, tables=tabs, mtos=mt, tail_rows=1))
    plan = _check(prob, pt, jobs)
    assert len(plan.levels) == 3                      # root, one node per t1 that at least two jobs share, the jobs
    root, mid, leaves = plan.levels
    assert root.n_traj == 1 and root.snap_steps.tolist() == [2, 5, 9, 14]
    assert mid.n_traj == 2 and mid.step0.tolist() == [2, 5] and mid.snap_cnt.tolist() == [3, 2]
    depth = plan.depth_of_job.reshape(-1)
    # t2 == t1: both operators at one step -> straight from the root; t2 > t1: from the t1 node
    assert sorted(depth.tolist()) == [0] * 5 + [1] * 5
    assert np.all(leaves.n_steps[depth == 1] == 15) and plan.n_slots == 4 + 5
    # the work: one-level forking would run every job from t1
    one_level = sum(int(round((j.t_end - j.mtos[0].time) / 0.1)) for j in jobs)
    assert int(leaves.n_steps.sum() + mid.n_steps.sum()) < one_level


def test_forks_respect_kept_rows_and_before_operators():
    """A job whose kept rows reach back before its second operator forks only at the first; 'applyBefore' operators
    and operators sharing a step keep their order; unforked execution agrees."""
    prob = biexciton_problem()
    pt = synthetic_pt(5, len(prob.cls_keys), kind="unitary", scale=0.999)
    p = ChirpedPulse(tau_0=0.4, e_start=-2.0, alpha=0, t0=1.0, e0=4.0)
    tabs = make_tables([p], 0.0, 4.0, 0.25)
    L = lambda t, before=False: _mto(prob, "|1><3|_4", "_left", t, before)
    R = lambda t, before=False: _mto(prob, "|3><1|_4", "_right", t, before)
    X = lambda t: _mto(prob, "|0><1|_4", "_left", t)
    jobs = [Job(0.0, 3.0, 0.25, tables=tabs, mtos=L(0.5) + R(0.5) + X(1.5), tail_rows=3),
            Job(0.0, 3.5, 0.25, tables=tabs, mtos=L(0.5) + R(0.5) + X(2.0), tail_rows=0),       # all rows
            Job(0.0, 3.0, 0.25, tables=tabs, mtos=L(0.5) + R(0.5) + X(1.0, ), tail_rows=10),    # rows before X
            Job(0.0, 3.0, 0.25, tables=tabs, mtos=R(0.5) + L(0.5) + X(1.0), tail_rows=2),       # same product, other order
            Job(0.0, 3.0, 0.25, tables=tabs, mtos=L(0.5, True) + X(2.0), tail_rows=2),
            Job(0.0, 3.0, 0.25, tables=tabs, mtos=L(0.5, True) + X(2.5), tail_rows=1),
            Job(0.0, 2.0, 0.25, tables=tabs),
            Job(0.0, 3.0, 0.25, tables=tabs, mtos=L(0.0) + X(1.0))]                             # operator at the start
    plan = _check(prob, pt, jobs)
    assert plan.depth_of_job.tolist() == [1, 0, 0, 1, 1, 1, -1, -1]
    assert len(plan.copies) == 1 and plan.copies[0].tolist() == [1, 2, 0, 0]
    _check(prob, pt, jobs, fork=False)


def test_truncated_drive_tables_and_groups():
    """Jobs that see only their own part of a shared drive table (Job.table_len), a second group with another
    start time and per-job initial states."""
    prob = tls_problem()
    pt = synthetic_pt(4, len(prob.cls_keys), kind="unitary", scale=0.999)
    p = ChirpedPulse(tau_0=0.5, e_start=0.5, alpha=0, t0=2.0, e0=3.0)
    tabs = make_tables([p], 0.0, 4.0, 0.1)
    tabs2 = make_tables([p], 0.0, 4.0, 0.1)
    M = lambda t: _mto(prob, "|0><1|_2", "_left", t)
    jobs = []
    for t_end, t_m in ((1.5, 1.0), (2.0, 1.0), (2.0, 1.9), (2.5, 1.2), (2.0, 2.0)):
        jb = Job(0.0, t_end, 0.1, tables=tabs, mtos=M(t_m), tail_rows=4)
        jb.table_len = len(np.arange(0.0, t_end, 0.1))
        jobs.append(jb)
    rho = np.array([0.3, 0.2 - 0.1j, 0.2 + 0.1j, 0.7])
    jobs += [Job(1.0, 3.0, 0.1, tables=tabs2, mtos=M(2.0)), Job(1.0, 3.0, 0.1, tables=tabs2, mtos=M(2.5)),
             Job(0.0, 2.0, 0.1, tables=tabs, mtos=M(1.0), rho0=rho), Job(0.0, 2.0, 0.1, tables=tabs, mtos=M(1.5), rho0=rho)]
    plan = _check(prob, pt, jobs)
    # the operator in a job's clamped last rows (1.9 of 2.0; 2.0 of 2.0) is not a fork point
    assert plan.depth_of_job.tolist() == [0, 0, -1, 0, -1, 0, 0, 0, 0]
    assert plan.levels[0].n_traj == 3


def test_planning_a_256_x_256_triangular_sweep_is_array_work():
    """32 896 jobs with three operators each planned from arrays: milliseconds, not a Python loop per job."""
    n, tb_steps = 256, 300
    i, j = np.triu_indices(n)
    J = len(i)
    t1, t2 = 1 + i, 1 + j
    diag = t1 == t2
    ev_step = np.stack([t1, np.where(diag, t1 + tb_steps, t2), np.where(diag, planner.BIG, t1 + tb_steps)], axis=1)
    ev_sb = np.full((J, 3), -1)
    ev_sa = np.stack([np.where(diag, 3, 0), np.where(diag, 2, 1), np.where(diag, -1, 2)], axis=1)
    z = np.zeros(J, dtype=np.int32)
    arr = planner.JobArrays(dt=0.1, t0=0.0, packed=np.zeros((1, 3, 8), complex), grid=(0.0, 0.1), mats=[None] * 4,
                            rho0s=np.zeros((1, 4), complex), set_id=z, shift=z, r0=z,
                            n_steps=(t2 + tb_steps).astype(np.int32), tail=z + 1, clamp=z,
                            ev_step=ev_step.astype(np.int32), ev_sb=ev_sb.astype(np.int32), ev_sa=ev_sa.astype(np.int32),
                            n_ev=np.where(diag, 2, 3).astype(np.int32), n_fork=np.where(diag, 2, 3).astype(np.int32))
    planner.plan_levels(arr, 2)
    t0 = time.perf_counter()
    plan = planner.plan_levels(arr, 2)
    ms = (time.perf_counter() - t0) * 1e3
    root, mid, leaves = plan.levels
    assert root.n_traj == 1 and mid.n_traj == n - 2 and leaves.n_traj == J
    assert int((plan.depth_of_job == 1).sum()) == J - n - 1 and plan.n_slots == n + (J - n - 1)
    assert np.sum(leaves.n_steps != tb_steps) == 1    # every job but one runs exactly one time bin
    assert ms < 50.0, ms


def test_sweep_arrays_equal_job_arrays():
    """The array route into the planner builds the same table as the per-job route: merged operators at one step
    (file order, 'before' separately), absent operators, clamp rows from per-job table lengths."""
    prob = biexciton_problem()
    p = ChirpedPulse(tau_0=0.4, e_start=-2.0, alpha=0, t0=1.0, e0=4.0)
    tabs = make_tables([p], 0.0, 6.0, 0.25)
    tmpl = [{"operator": "|1><3|_4", "applyFrom": "_left", "applyBefore": "false"},
            {"operator": "|3><1|_4", "applyFrom": "_right", "applyBefore": "true"},
            {"operator": "|0><1|_4", "applyFrom": "_left", "applyBefore": "false"}]
    times = np.array([[0.5, 0.5, 1.5], [1.0, 0.5, 1.0], [2.0, 2.0, 2.0], [0.0, np.nan, 3.0], [np.nan, np.nan, np.nan],
                      [2.75, 1.0, 3.0], [3.0, 2.75, 1.0], [3.0, np.nan, 1.0], [np.nan, 2.5, np.nan]])
    t_end = np.array([3.0, 3.5, 2.0, 3.0, 1.0, 3.0, 3.0, 3.0, 2.5])
    tails = np.array([3, 0, 1, 2, 0, 4, 1, 0, 2])
    lens = np.array([len(np.arange(0.0, te, 0.25)) for te in t_end])
    jobs = []
    for row, te, tl, ln in zip(times, t_end, tails, lens):
        mt = [dict(m, time=float(t)) for m, t in zip(tmpl, row) if not np.isnan(t)]
        jb = Job(0.0, float(te), 0.25, tables=tabs, mtos=prob.parse_mtos(mt), tail_rows=int(tl))
        jb.table_len = int(ln)
        jobs.append(jb)
    a = planner.arrays_from_jobs(prob, jobs)
    parsed = prob.parse_mtos([dict(m, time=0.0) for m in tmpl])
    b = planner.arrays_from_sweep(prob, dt=0.25, t_start=0.0, t_end=t_end, superops=[m.superop for m in parsed],
                                  before=[m.before for m in parsed], mto_times=times, tails=tails, tables=tabs,
                                  table_len=lens)
    E = a.ev_step.shape[1]
    assert np.array_equal(a.n_ev, b.n_ev) and np.array_equal(a.n_fork, b.n_fork) and np.array_equal(a.clamp, b.clamp)
    assert np.array_equal(a.ev_step, b.ev_step[:, :E]) and np.all(b.ev_step[:, E:] == planner.BIG)
    for x, y in ((a.ev_sb, b.ev_sb), (a.ev_sa, b.ev_sa)):
        for j in range(len(jobs)):
            for e in range(int(a.n_ev[j])):
                assert (x[j, e] < 0) == (y[j, e] < 0)
                if x[j, e] >= 0:
                    assert np.array_equal(a.mats[x[j, e]], b.mats[y[j, e]])
    with pytest.raises(ValueError):
        planner.arrays_from_sweep(prob, dt=0.25, t_start=0.0, t_end=t_end, superops=[m.superop for m in parsed],
                                  before=[m.before for m in parsed], mto_times=times + 0.1, tails=tails, tables=tabs)
    pt = synthetic_pt(5, len(prob.cls_keys), kind="unitary", scale=0.999)
    got = run_plan(prob, pt, jobs, planner.plan_levels(b, prob.n_out))
    for jb, g in zip(jobs, got):
        want = oracle.propagate(prob, pt, jb)
        assert np.abs(g - (want[:, -jb.tail_rows:] if jb.tail_rows else want)).max() < 1e-11
