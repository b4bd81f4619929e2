"""GPU parity of the split-K cluster kernel (csrc/splitk_kernel.cu): a cluster of C CTAs holds chi_pad/C bond columns
each of G trajectories; partial PT products are reduce-scattered through distributed shared memory.  Same oracle,
same 1e-10 tolerance as the tile kernel; the library's record of the launched kernel is asserted."""
import numpy as np
import pytest

import oracle
from helpers import biexciton_problem, make_tables, sixls_problem, sweep_jobs, tls_problem
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import MTO
from pyaceqd_b200.process_tensor import synthetic_growing_pt, synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _compare(engine, prob, pt, jobs, **kw):
    got = engine.run_jobs(prob, pt, jobs, kernel="splitk", **kw)
    assert engine.last_kernels()["step"].startswith("k_step_splitk"), engine.last_kernels()
    worst = 0.0
    for g, jb in zip(got, jobs):
        ref = oracle.propagate(prob, pt, jb)
        if jb.tail_rows:
            ref = ref[:, -jb.tail_rows:]
        assert g.shape == ref.shape
        worst = max(worst, float(np.abs(g - ref).max()))
    assert worst < TOL, f"max abs deviation {worst:.3e} ({kw})"
    return got


@pytest.mark.parametrize("chi", [8, 20, 40, 64, 72, 128])
@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_tls_sweeps_every_cluster_size(engine, chi, cluster):
    prob = tls_problem()
    pt = synthetic_pt(chi, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs = sweep_jobs(5, 7, t_end=3.0)        # 35 trajectories: ragged last tile
    for G in (1, 3, 8, 16):
        _compare(engine, prob, pt, jobs, cluster=cluster, tile_T=G)
        assert "G=%d cluster=%d" % (G, cluster) in engine.last_kernels()["step"]


@pytest.mark.parametrize("chi", [136, 200, 256])
def test_panelled_passes_above_128_bond_columns(engine, chi):
    """chi_pad > 128: every GEMM pass is cut into two panels of 128 output columns (panel-ordered PT copy)."""
    prob = tls_problem()
    pt = synthetic_pt(chi, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs = sweep_jobs(3, 5, t_end=2.0)
    for cluster, G in ((2, 4), (4, 8), (8, 16), (8, 5)):
        _compare(engine, prob, pt, jobs, cluster=cluster, tile_T=G)
        assert "panels=2" in engine.last_kernels()["step"]


def test_growing_pt_ragged_lengths_and_mtos(engine):
    from pyaceqd_b200.opparser import parse_operator
    prob = tls_problem()
    pt = synthetic_growing_pt(24, len(prob.cls_keys), n_initial=5, n_repeat=3)
    p = ChirpedPulse(tau_0=1, e_start=0.5, alpha=0, t0=3, e0=2)
    s = parse_operator("|0><1|_2", 2)
    jobs = []
    for te in (0.0, 0.1, 0.3, 1.0, 2.7, 5.0, 3.3):
        mt = [MTO(prob.mto_superop(s, ""), 0.0, False)]
        if te >= 1.0:
            mt += [MTO(prob.mto_superop(s.conj().T, "_left"), 0.5, True), MTO(prob.mto_superop(s, "_right"), 0.5, False),
                   MTO(prob.mto_superop(s.conj().T, "_right"), te, True)]
        jobs.append(Job(0.0, te, 0.1, tables=make_tables([p], 0.0, max(te, 0.1), 0.1), mtos=mt))
    for cluster, G in ((2, 4), (4, 3), (8, 8)):
        _compare(engine, prob, pt, jobs, cluster=cluster, tile_T=G, fork=False)


@pytest.mark.parametrize("cluster,G", [(2, 2), (4, 4), (4, 8), (8, 8), (8, 16)])
def test_biexciton_g2_fork_tails_and_snapshots(engine, cluster, G):
    """Forked G2-style batch: branches on the split-K kernel start from snapshots written by the tile kernel's trunk
    (closures travel with the snapshots) -- and, with trunk_kernel="splitk", by the split-K kernel itself."""
    prob = biexciton_problem(outputs=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)", "|0><3|_4"])
    pt = synthetic_pt(40, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    dt, tau_max = 0.25, 3.0
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=2.0, e0=4.0, polar_x=0.8)
    tabs = make_tables([p], 0.0, 12.0, dt)

    def jobs_of(tail):
        out = []
        for i in range(11):
            t1 = 2 * i * dt
            mt = prob.parse_mtos([{"operator": "|3><1|_4", "applyFrom": "_right", "time": t1},
                                  {"operator": "|1><3|_4", "applyFrom": "_left", "time": t1}])
            out.append(Job(0.0, t1 + tau_max, dt, tables=tabs, mtos=mt, tail_rows=tail))
        return out

    engine.record_timings = True
    engine.timing_log.clear()
    try:
        _compare(engine, prob, pt, jobs_of(0), cluster=cluster, tile_T=G)
        kinds = {l["kind"]: l["step_kernel"] for l in engine.timing_log}
        assert kinds["trunk"].startswith("k_step_dmma") and kinds["main"].startswith("k_step_splitk"), kinds
        engine.timing_log.clear()
        _compare(engine, prob, pt, jobs_of(5), cluster=cluster, tile_T=G, trunk_kernel="splitk")
        kinds = {l["kind"]: l["step_kernel"] for l in engine.timing_log}
        assert kinds["trunk"].startswith("k_step_splitk"), kinds
    finally:
        engine.record_timings = False
    _compare(engine, prob, pt, jobs_of(0), cluster=cluster, tile_T=G, fork=False)


def test_sixlevel_and_planner_choice(engine):
    six = sixls_problem()
    pt6 = synthetic_pt(24, len(six.cls_keys), kind="unitary", scale=0.999)
    p6 = ChirpedPulse(tau_0=1.0, e_start=-1.0, alpha=0, t0=2.0, e0=3.0, polar_x=0.7)
    jobs6 = [Job(0.0, 3.0, 0.1, tables=make_tables([ChirpedPulse(tau_0=1.0, e_start=-1.0, alpha=0, t0=2.0, e0=a, polar_x=0.7)],
                                                    0.0, 3.0, 0.1)) for a in (1.0, 2.0, 3.0, 4.0, 5.0)]
    _compare(engine, six, pt6, jobs6)                       # planner's own (G, C)
    _compare(engine, six, pt6, jobs6, cluster=4, tile_T=3)
    # the planner keeps the cfg3 branch launch on small clusters (the exchange cost grows with the cluster size:
    # measured 14.4 ms on (8, 4) and (4, 2) against 35.8 ms on (16, 8), profiles/r05j_gc_sweep.txt)
    bx = biexciton_problem()
    ptb = synthetic_pt(128, len(bx.cls_keys), kind="unitary", scale=0.999)
    G, C, _ = engine._splitk_tile(bx, ptb, 256)
    assert C in (2, 4) and G >= 4, (G, C)


def test_shapes_too_large_for_a_pt_ring_run_without_one_or_on_the_cluster_kernel(engine):
    """NL = 36 at chi = 256: ONE trajectory (36 x 260 x 16 B = 150 KB) leaves no room for a shared-memory ring of PT
    chunks.  The tile kernel then reads the PT fragments from global memory / L2 ("pt=global"); the split-K cluster
    kernel (bond columns spread over a cluster) stays available for the same job."""
    six = sixls_problem()
    pt = synthetic_pt(256, len(six.cls_keys), kind="unitary", scale=0.999)
    assert engine.max_tile_ring(six.NL, 256) == 0 and engine.max_tile(six.NL, 256) == 1
    p6 = ChirpedPulse(tau_0=1.0, e_start=-1.0, alpha=0, t0=1.0, e0=3.0, polar_x=0.7)
    jobs = [Job(0.0, 1.5, 0.1, tables=make_tables([p6], 0.0, 1.5, 0.1)) for _ in range(2)]
    for kernel, want in (("auto", "pt=global"), ("splitk", "k_step_splitk")):
        got = engine.run_jobs(six, pt, jobs, kernel=kernel)
        assert want in engine.last_kernels()["step"], engine.last_kernels()
        for g, jb in zip(got, jobs):
            assert np.abs(g - oracle.propagate(six, pt, jb)).max() < TOL
