"""GPU parity: the CUDA path (through the C ABI, ctypes) against the CPU oracle on identical
inputs.  Tolerance: 1e-10 absolute on every output (north star asks <= 1e-8 in FP64)."""
import numpy as np
import pytest

import oracle
from helpers import biexciton_problem, make_tables, sixls_problem, sweep_jobs, tls_problem
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import MTO
from pyaceqd_b200.process_tensor import synthetic_growing_pt, synthetic_pt, trivial_pt
from pyaceqd_b200.pulses import ChirpedPulse

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _compare(engine, prob, pt, jobs, kernel, **kw):
    got = engine.run_jobs(prob, pt, jobs, kernel=kernel, **kw)
    worst = 0.0
    for g, jb in zip(got, jobs):
        ref = oracle.propagate(prob, pt, jb)
        assert g.shape == ref.shape
        worst = max(worst, float(np.abs(g - ref).max()))
    assert worst < TOL, f"max abs deviation {worst:.3e} (kernel={kernel})"
    return worst


def test_dmma_fragment_layout_via_expm(engine):
    """Device expm (Taylor scaling-and-squaring) vs scipy Pade."""
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    for n in (4, 9, 16, 25, 36, 49, 64):
        a = (rng.standard_normal((7, n, n)) + 1j * rng.standard_normal((7, n, n))) * rng.uniform(0.01, 3.0, (7, 1, 1))
        got = engine.expm(a)
        for i in range(7):
            ref = expm(a[i])
            assert np.abs(got[i] - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("kernel", ["check", "dmma"])
def test_tls_no_phonons(engine, kernel):
    prob = tls_problem(phonons=False)
    p = ChirpedPulse(tau_0=3, e_start=0, alpha=0, t0=12, e0=1)
    jobs = [Job(0.0, 30.0, 0.1, tables=make_tables([p], 0.0, 30.0, 0.1))]
    _compare(engine, prob, trivial_pt(len(prob.cls_keys)), jobs, kernel)


@pytest.mark.parametrize("kernel", ["check", "dmma"])
@pytest.mark.parametrize("chi", [8, 20, 64, 128])
def test_tls_synthetic_pt_sweep(engine, kernel, chi):
    prob = tls_problem()
    pt = synthetic_pt(chi, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs = sweep_jobs(6, 7, t_end=4.0 if chi > 32 else 8.0)
    _compare(engine, prob, pt, jobs, kernel)


@pytest.mark.parametrize("kernel", ["check", "dmma"])
def test_growing_pt_and_ragged_lengths(engine, kernel):
    prob = tls_problem()
    pt = synthetic_growing_pt(24, len(prob.cls_keys), n_initial=5, n_repeat=3)
    p = ChirpedPulse(tau_0=1, e_start=0.5, alpha=0, t0=3, e0=2)
    jobs = [Job(0.0, te, 0.1, tables=make_tables([p], 0.0, te, 0.1)) for te in (0.0, 0.1, 0.3, 1.0, 2.7, 5.0)]
    _compare(engine, prob, pt, jobs, kernel)


@pytest.mark.parametrize("kernel", ["check", "dmma"])
def test_biexciton_mto_fork_equals_unforked_oracle(engine, kernel):
    """G2-style grid: every job shares the drive, MTOs at t1_i; the engine forks
    from one trunk, the oracle runs every trajectory from scratch."""
    from pyaceqd_b200.opparser import parse_operator
    prob = biexciton_problem(outputs=["|1><1|_4", "(|1><0|_4*|1><1|_4*|0><1|_4)", "|0><3|_4"])
    pt = synthetic_pt(32, len(prob.cls_keys), n_slices=3, kind="unitary", scale=0.999)
    dt, tau_max = 0.25, 3.0
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
    tabs = make_tables([p], 0.0, 12.0, dt)
    a = parse_operator("|1><0|_4", 4)
    c = parse_operator("|0><1|_4", 4)
    jobs = []
    for t1 in np.arange(0.0, 6.0, 0.75):
        mtos = [MTO(prob.mto_superop(a, "_right"), float(t1), False),
                MTO(prob.mto_superop(c, "_left"), float(t1), False)]
        jobs.append(Job(0.0, float(t1 + tau_max), dt, tables=tabs, mtos=mtos))
    w1 = _compare(engine, prob, pt, jobs, kernel, fork=True)
    w2 = _compare(engine, prob, pt, jobs, kernel, fork=False)
    assert max(w1, w2) < TOL


@pytest.mark.parametrize("kernel", ["check", "dmma"])
def test_mto_before_sandwich_and_multiple_times(engine, kernel):
    from pyaceqd_b200.opparser import parse_operator
    prob = tls_problem()
    pt = synthetic_pt(16, len(prob.cls_keys), kind="unitary")
    p = ChirpedPulse(tau_0=1.0, e_start=0.3, alpha=0, t0=2.0, e0=1.5)
    tabs = make_tables([p], 0.0, 6.0, 0.1)
    s = parse_operator("|0><1|_2", 2)
    jobs = [Job(0.0, 6.0, 0.1, tables=tabs, mtos=[
        MTO(prob.mto_superop(s, ""), 0.0, False), MTO(prob.mto_superop(s.conj().T, "_left"), 1.5, True),
        MTO(prob.mto_superop(s, "_right"), 1.5, False), MTO(prob.mto_superop(s.conj().T, "_right"), 6.0, True)])]
    _compare(engine, prob, pt, jobs, kernel)


@pytest.mark.parametrize("kernel", ["check", "dmma"])
def test_sixlevel_nl36(engine, kernel):
    prob = sixls_problem()
    pt = synthetic_pt(40, len(prob.cls_keys), kind="unitary")
    p1 = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=5.0, polar_x=0.7)
    jobs = [Job(0.0, 2.0, 0.1, tables=make_tables([p1], 0.0, 2.0, 0.1)) for _ in range(3)]
    _compare(engine, prob, pt, jobs, kernel)


def test_tile_sizes_agree(engine):
    prob = tls_problem()
    pt = synthetic_pt(32, len(prob.cls_keys))
    jobs = sweep_jobs(5, 5, t_end=3.0)
    ref = engine.run_jobs(prob, pt, jobs, kernel="check")
    for T in (1, 2, 4, 8, 16):
        got = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=T)
        assert max(np.abs(g - r).max() for g, r in zip(got, ref)) < 1e-12, T


# ------------------------------------------------------------------ thread-block clusters (row exchange over DSMEM)
def _g2_jobs(prob, n_t=6, dt=0.25, tau_max=3.0, opA="|3><1|_4", opC="|1><3|_4", tail=0):
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=2.0, e0=4.0, polar_x=0.8)
    tabs = make_tables([p], 0.0, n_t * dt * 2 + tau_max + 1, dt)
    jobs = []
    for i in range(n_t):
        t1 = 2 * i * dt
        mt = prob.parse_mtos([{"operator": opA, "applyFrom": "_right", "time": t1},
                              {"operator": opC, "applyFrom": "_left", "time": t1}])
        jobs.append(Job(0.0, t1 + tau_max, dt, tables=tabs, mtos=mt, tail_rows=tail))
    return jobs


@pytest.mark.parametrize("cluster,tile_T", [(2, 1), (2, 2), (2, 4), (4, 1), (4, 2), (4, 4), (8, 1), (8, 2), (16, 1), (16, 2)])
def test_cluster_biexciton_fork_and_tails(engine, cluster, tile_T):
    """Tile shared by a cluster of CTAs: forked G2-style batch (trunk with snapshots + branches)."""
    prob = biexciton_problem(outputs=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)", "|0><3|_4"])
    pt = synthetic_pt(40, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs = _g2_jobs(prob)
    _compare(engine, prob, pt, jobs, "dmma", cluster=cluster, tile_T=tile_T)
    assert engine.last_kernels()["step"] == "k_step_dmma<1,4> T=%d cluster=%d segments=0" % (tile_T, cluster)
    assert engine.last_kernels()["opbuild"] == "k_opbuild_dmma<2>"
    tails = engine.run_jobs(prob, pt, _g2_jobs(prob, tail=5), cluster=cluster, tile_T=tile_T)
    full = engine.run_jobs(prob, pt, jobs, cluster=1, tile_T=tile_T)
    assert max(np.abs(f[:, -5:] - t).max() for f, t in zip(full, tails)) < 1e-12


def test_cluster16_falls_back_to_8_where_it_cannot_be_scheduled(engine, monkeypatch):
    """16 CTAs are beyond the portable cluster size; the library asks cudaOccupancyMaxActiveClusters and runs the tile
    on 8 CTAs where such a cluster cannot be placed (simulated here)."""
    prob = biexciton_problem(outputs=["|1><1|_4", "|0><3|_4"])
    pt = synthetic_pt(40, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs = _g2_jobs(prob)
    monkeypatch.setenv("ACEQD_CLUSTER16_UNSCHEDULABLE", "1")
    _compare(engine, prob, pt, jobs, "dmma", cluster=16, tile_T=1)
    assert engine.last_kernels()["step"] == "k_step_dmma<1,4> T=1 cluster=8 segments=0"
    monkeypatch.delenv("ACEQD_CLUSTER16_UNSCHEDULABLE")
    _compare(engine, prob, pt, jobs, "dmma", cluster=16, tile_T=1)
    assert engine.last_kernels()["step"] == "k_step_dmma<1,4> T=1 cluster=16 segments=0"


@pytest.mark.parametrize("cluster", [2, 4])
def test_cluster_tls_sweep_and_sixlevel(engine, cluster):
    prob = tls_problem()
    pt = synthetic_pt(64, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    _compare(engine, prob, pt, sweep_jobs(5, 5, t_end=3.0), "dmma", cluster=cluster, tile_T=8)
    pt = synthetic_growing_pt(24, len(prob.cls_keys), n_initial=5, n_repeat=3)
    p = ChirpedPulse(tau_0=1, e_start=0.5, alpha=0, t0=3, e0=2)
    jobs = [Job(0.0, te, 0.1, tables=make_tables([p], 0.0, te, 0.1)) for te in (0.0, 0.1, 0.3, 1.0, 2.7, 5.0)]
    _compare(engine, prob, pt, jobs, "dmma", cluster=cluster, tile_T=4)         # ragged lengths, growing bond
    six = sixls_problem()
    pt6 = synthetic_pt(24, len(six.cls_keys), kind="unitary", scale=0.999)
    p6 = ChirpedPulse(tau_0=1.0, e_start=-1.0, alpha=0, t0=2.0, e0=3.0, polar_x=0.7)
    jobs6 = [Job(0.0, 3.0, 0.1, tables=make_tables([p6], 0.0, 3.0, 0.1)) for _ in range(3)]
    _compare(engine, six, pt6, jobs6, "dmma", cluster=cluster, tile_T=2)


@pytest.mark.parametrize("cluster", [1, 2])
def test_large_tile_single_buffered_operators(engine, cluster):
    """NL=16, chi=128, T=4 leaves room for only one staging buffer of the per-row operators (refilled
    during the PT GEMM) -- the cfg3 configuration."""
    prob = biexciton_problem(outputs=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)"])
    pt = synthetic_pt(128, len(prob.cls_keys), kind="unitary", scale=0.999)
    _compare(engine, prob, pt, _g2_jobs(prob, n_t=8, tau_max=2.0), "dmma", cluster=cluster, tile_T=4)


def test_planner_picks_clusters_for_small_batches(engine):
    """256 biexciton branches at chi=128 cannot fill 148 SMs with one CTA per tile."""
    prob = biexciton_problem()
    pt = synthetic_pt(128, len(prob.cls_keys), kind="unitary", scale=0.999)
    t_max = engine.max_tile(prob.NL, 128)
    T, C = engine._tile_and_cluster(prob, pt, 256, t_max)
    assert (T, C) == (4, 2)
    assert engine._tile_and_cluster(prob, pt, 1, t_max) == (1, 16)      # the lone trunk: one GEMM pass per CTA (9 classes)
    assert engine._tile_and_cluster(prob, pt, 9, t_max)[1] <= 8         # 16-CTA clusters only while they are all resident
    tls = tls_problem()
    ptt = synthetic_pt(128, len(tls.cls_keys), kind="unitary", scale=0.999)
    assert engine._tile_and_cluster(tls, ptt, 4096, engine.max_tile(4, 128)) == (16, 1)


def test_more_tiles_than_sms_segments_waves_and_copy_overlap(engine, monkeypatch):
    """More tiles than SMs on the tile kernel (chi = 40 > 32, so the small-bond kernel cannot take the batch), through
    host buffers: (a) the default segment schedule (tiles cut into one piece of steps per SM, bond states handed over
    through HBM), (b) ACEQD_SEGMENTS=0: the partial wave first, operators of the later waves built on a side stream,
    (c) pageable outputs: finished waves copied while the last wave runs.  Results must not depend on any of it."""
    prob = tls_problem()
    pt = synthetic_pt(40, len(prob.cls_keys), kind="unitary", scale=0.999)
    jobs = sweep_jobs(20, 30, t_end=2.0)                          # 600 trajectories
    whole = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=16, cluster=1)     # 38 tiles: single launch
    assert engine.last_kernels()["step"] == "k_step_dmma<1,1> T=16 cluster=1 segments=0"
    for k in (0, 147 * 2, 148 * 2, 599):                          # around the wave boundary
        assert np.abs(whole[k] - oracle.propagate(prob, pt, jobs[k])).max() < TOL
    for env, seg in ((None, 1), ("0", 0)):
        if env is None:
            monkeypatch.delenv("ACEQD_SEGMENTS", raising=False)
        else:
            monkeypatch.setenv("ACEQD_SEGMENTS", env)
        for nzc in ("0", "1"):
            monkeypatch.setenv("ACEQD_NO_ZEROCOPY", nzc)
            split = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=2, cluster=1)      # 300 tiles > 148 SMs
            assert engine.last_kernels()["step"] == "k_step_dmma<1,1> T=2 cluster=1 segments=%d" % seg
            assert max(np.abs(a - b).max() for a, b in zip(split, whole)) < 1e-12, (env, nzc)
    monkeypatch.delenv("ACEQD_SEGMENTS", raising=False)
    monkeypatch.delenv("ACEQD_NO_ZEROCOPY", raising=False)


# ------------------------------------------------------------------ small-bond kernel (one warp per 8 trajectories)
@pytest.mark.parametrize("chi", [5, 8, 16, 20, 32])
def test_small_bond_kernel_matches_oracle_and_tile_kernel(engine, chi):
    """NL = 4, chi_pad <= 32 on k_step_small<chi_pad/8> (PT resident in shared memory, warp-private state), forced with
    kernel="small" and PROVEN by the library's record of the kernel it launched.  Ragged lengths and octets,
    multi-slice PTs, MTO rows (override entries), tails, more outputs than lanes of a quad, and a forked batch whose
    branches start from snapshots written by the tile kernel's trunk."""
    from pyaceqd_b200.opparser import parse_operator
    prob = tls_problem()
    pt = synthetic_pt(chi, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    nt = -(-chi // 8)
    small = "k_step_small<%d>" % nt

    def ran_small():
        assert engine.last_kernels()["step"].startswith(small), engine.last_kernels()

    jobs = sweep_jobs(7, 9, t_end=6.0)                      # 63 trajectories: octets with a ragged tail
    got = engine.run_jobs(prob, pt, jobs, kernel="small")
    ran_small()
    for k in range(0, len(jobs), 5):
        assert np.abs(got[k] - oracle.propagate(prob, pt, jobs[k])).max() < TOL
    ref = engine.run_jobs(prob, pt, jobs, kernel="tile")
    assert engine.last_kernels()["step"].startswith("k_step_dmma<1,1>")
    assert max(np.abs(a - b).max() for a, b in zip(got, ref)) < 1e-12
    # ragged lengths + MTOs at several times (override rows)
    p = ChirpedPulse(tau_0=1.0, e_start=0.3, alpha=0, t0=2.0, e0=1.5)
    s = parse_operator("|0><1|_2", 2)
    rag = []
    for te in (0.0, 0.1, 0.4, 1.0, 2.7, 3.3, 4.1, 5.0, 5.0, 2.2, 0.9):
        tabs = make_tables([p], 0.0, max(te, 0.1), 0.1)
        mt = [MTO(prob.mto_superop(s, ""), 0.0, False)]
        if te >= 1.0:
            mt += [MTO(prob.mto_superop(s.conj().T, "_left"), 0.5, True), MTO(prob.mto_superop(s, "_right"), 0.5, False)]
        rag.append(Job(0.0, te, 0.1, tables=tabs, mtos=mt))
    _compare(engine, prob, pt, rag, "small", fork=False)
    ran_small()
    # forked G1-style batch: trunk (tile kernel, snapshots) + branches (small kernel, init from snapshots)
    tabs = make_tables([p], 0.0, 12.0, 0.1)
    fj = []
    for t1 in np.arange(0.5, 5.0, 0.5):
        mt = prob.parse_mtos([{"operator": "|0><1|_2", "applyFrom": "_left", "time": float(t1)}])
        fj.append(Job(0.0, float(t1 + 3.0), 0.1, tables=tabs, mtos=mt))
    engine.record_timings = True
    engine.timing_log.clear()
    try:
        _compare(engine, prob, pt, fj, "small", fork=True)
        kinds = {l["kind"]: l["step_kernel"] for l in engine.timing_log}
    finally:
        engine.record_timings = False
    assert kinds["trunk"].startswith("k_step_dmma") and kinds["main"].startswith(small), kinds
    # more output functionals than lanes of a quad (full density matrix + two products)
    prob6 = tls_problem(outputs=["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2", "(|1><0|_2*|0><1|_2)", "|0><0|_2+|1><1|_2"])
    _compare(engine, prob6, pt, sweep_jobs(3, 5, t_end=3.0), "small")
    ran_small()
    tails = engine.run_jobs(prob, pt, [Job(j.t_start, j.t_end, j.dt, tables=j.tables, mtos=j.mtos, tail_rows=11)
                                       for j in fj], kernel="small")
    ran_small()
    full = engine.run_jobs(prob, pt, fj, kernel="tile")
    assert max(np.abs(f[:, -11:] - t).max() for f, t in zip(full, tails)) < 1e-12


def test_kernel_choice_of_the_library(engine):
    """kernel 0 takes k_step_small only for batches that can fill the sub-partitions; small batches stay on tiles
    (shared by clusters).  Ineligible batches forced onto the small kernel fail loudly."""
    from pyaceqd_b200.engine import SMALL_MIN_TRAJ, EngineError
    prob = tls_problem()
    pt = synthetic_pt(16, len(prob.cls_keys), kind="unitary", scale=0.999)
    jobs = sweep_jobs(6, 6, t_end=1.0)
    engine.run_jobs(prob, pt, jobs)
    assert engine.last_kernels()["step"].startswith("k_step_dmma<1,1>")
    p = ChirpedPulse(tau_0=1.0, e_start=0.3, alpha=0, t0=0.5, e0=1.5)
    tabs = make_tables([p], 0.0, 1.0, 0.1)
    many = [Job(0.0, 1.0, 0.1, tables=tabs) for _ in range(SMALL_MIN_TRAJ)]
    out = engine.run_jobs(prob, pt, many)
    assert engine.last_kernels()["step"].startswith("k_step_small<2>")
    assert engine.last_kernels()["opbuild"] == "k_opbuild_reg<4>"
    assert np.abs(out[-1] - oracle.propagate(prob, pt, many[-1])).max() < TOL
    with pytest.raises(EngineError):
        engine.run_jobs(prob, synthetic_pt(40, len(prob.cls_keys), kind="unitary"), jobs, kernel="small")
