"""GPU parity at the shapes BASELINE.json names, on process tensors that KEEP the signal:

* cfg3  biexciton NL=16, chi=128: the full 256 branches x 256 steps G2 grid (trunk + forked branches), a sample of
        branches against the un-forked oracle;
* cfg4  six-level NL=36, chi=128 with the G2_reuse operator pattern (reference pol_entanglement/G2.py:243-299);
* cfg5  five-level dark model NL=25 (reference four_level_system/dark_model.py:34-55), chi=256, with the three
        multi-time operators of the timebin sweeps (reference timebin/twophoton_new.py:515-557);
* physical process tensors from the host builder (two-level threshold 1e-8 -> chi ~ 27; biexciton dt=0.5 -> chi ~ 116)
  pushed through the CUDA path.

Synthetic PTs are `kind="unitary"` (spectral radius ~1), so the 1e-10 absolute tolerance bites on every row; each
test also asserts that the compared signal is O(1e-2) or larger.  The library's record of the kernel it launched is
asserted where a test is about a particular kernel.
"""
import numpy as np
import pytest

import oracle
from helpers import biexciton_problem, make_tables, sixls_problem, tls_problem
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import build_problem
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse

pytestmark = pytest.mark.gpu
TOL = 1e-10


def fivels_problem(outputs=None):
    """darkmodel_new of the reference (four_level_system/dark_model.py:34-55) with radiative loss."""
    return build_problem(
        system_op=["-4.0*|4><4|_5", "-0.1*|3><3|_5"], boson_op="1*(|1><1|_5 + |2><2|_5 + |3><3|_5) + 2*|4><4|_5",
        initial="|0><0|_5", lindblad_ops=[["|0><1|_5", 0.01], ["|0><2|_5", 0.01], ["|1><4|_5", 0.01], ["|2><4|_5", 0.01]],
        interaction_ops=[["|1><0|_5", "x"], ["|4><1|_5", "x"], ["|3><0|_5", "y"], ["|4><3|_5", "y"]],
        output_ops=outputs or ["|0><0|_5", "|1><1|_5", "|4><4|_5", "|0><4|_5", "|1><3|_5"])


def _check(engine, prob, pt, jobs, sample, min_signal=1e-2, **kw):
    got = engine.run_jobs(prob, pt, jobs, **kw)
    worst, signal = 0.0, 0.0
    for i in sample:
        ref = oracle.propagate(prob, pt, jobs[i])
        if jobs[i].tail_rows:
            ref = ref[:, -jobs[i].tail_rows:]
        assert got[i].shape == ref.shape
        worst = max(worst, float(np.abs(got[i] - ref).max()))
        signal = max(signal, float(np.abs(ref[:, -1]).max()))
    assert signal > min_signal, f"the compared rows carry no signal ({signal:.1e})"
    assert worst < TOL, f"max abs deviation {worst:.3e}"
    return got


@pytest.mark.parametrize("kernel", ["dmma", "splitk"])
def test_cfg3_full_grid_256_branches_x_256_steps(engine, kernel):
    prob = biexciton_problem(outputs=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)"])
    pt = synthetic_pt(128, len(prob.cls_keys), kind="unitary", scale=0.999)
    dt, n_t = 0.25, 256
    tau_max = n_t * dt
    pulse = ChirpedPulse(tau_0=5.0, e_start=-2.0, alpha=0, t0=20.0, e0=5.0, polar_x=1.0)
    tabs = make_tables([pulse], 0.0, 2 * tau_max + 1.0, dt)
    jobs = []
    for i in range(n_t):
        t1 = i * dt
        mt = prob.parse_mtos([{"operator": "|3><1|_4", "applyFrom": "_right", "time": t1},
                              {"operator": "|1><3|_4", "applyFrom": "_left", "time": t1}])
        jobs.append(Job(0.0, t1 + tau_max, dt, tables=tabs, mtos=mt, tail_rows=n_t + 1))
    engine.record_timings = True
    engine.timing_log.clear()
    try:
        _check(engine, prob, pt, jobs, [0, 1, 37, 100, 128, 200, 254, 255], kernel=kernel)
        log = {l["kind"]: l for l in engine.timing_log}
    finally:
        engine.record_timings = False
    assert log["main"]["n_traj"] == n_t and log["trunk"]["n_traj"] == 1
    want = "k_step_splitk" if kernel == "splitk" else "k_step_dmma<2,4>"
    assert log["main"]["step_kernel"].startswith(want), log["main"]


@pytest.mark.parametrize("kernel", ["dmma", "splitk"])
def test_cfg4_sixlevel_nl36_chi128_g2_reuse_pattern(engine, kernel):
    prob = sixls_problem()
    pt = synthetic_pt(128, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    dt = 0.1
    p1 = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=1.5, e0=5.0, polar_x=0.7)
    tabs = make_tables([p1], 0.0, 6.0, dt)
    jobs = []
    for i in range(6):
        t1 = 0.3 * i
        mt = prob.parse_mtos([{"operator": "|0><1|_6", "applyFrom": "_left", "time": t1},
                              {"operator": "|2><0|_6", "applyFrom": "_right", "time": t1}])
        jobs.append(Job(0.0, t1 + 2.0, dt, tables=tabs, mtos=mt))
    _check(engine, prob, pt, jobs, range(len(jobs)), min_signal=1e-3, kernel=kernel)


@pytest.mark.parametrize("kernel,tile", [("dmma", 1), ("dmma", 2), ("dmma", None), ("splitk", None)])
def test_cfg5_fivelevel_nl25_chi256_three_mtos(engine, kernel, tile):
    """tile = 1: one trajectory per CTA and a shared-memory ring for the PT chunks; tile = 2: two trajectories fill
    shared memory, the PT fragments come from global memory / L2 (k_step_dmma<...,GPT>, launch name "pt=global")."""
    prob = fivels_problem()
    pt = synthetic_pt(256, len(prob.cls_keys), kind="unitary", scale=0.999)
    dt = 0.1
    p1 = ChirpedPulse(tau_0=0.5, e_start=-2.0, alpha=0, t0=1.0, e0=5.0, polar_x=0.8)
    tabs = make_tables([p1], 0.0, 5.0, dt)
    jobs = []
    for (a, b, c) in ((0.2, 0.5, 0.9), (0.2, 0.5, 1.4), (0.2, 0.8, 1.1), (0.4, 0.4, 1.0), (0.0, 0.6, 0.6), (0.3, 1.0, 1.2)):
        mt = prob.parse_mtos([{"operator": "|0><1|_5", "applyFrom": "_left", "time": a},
                              {"operator": "|0><1|_5", "applyFrom": "_left", "time": b},
                              {"operator": "|1><0|_5", "applyFrom": "_right", "time": c}])
        jobs.append(Job(0.0, c + 1.0, dt, tables=tabs, mtos=mt))
    engine.record_timings = True
    engine.timing_log.clear()
    try:
        _check(engine, prob, pt, jobs, range(len(jobs)), min_signal=1e-4, kernel=kernel, fork=tile is None, tile_T=tile)
        names = [l["step_kernel"] for l in engine.timing_log]
    finally:
        engine.record_timings = False
    if tile == 2:
        assert all("T=2" in n and "pt=global" in n for n in names), names
    if tile == 1:
        assert all("T=1" in n and "pt=global" not in n for n in names), names


def test_physical_pts_from_the_host_builder_on_the_gpu(engine):
    """a9: PTs made by pt_builder (QDPhonon spectral density, 4 K) -- not synthetic tensors -- through the CUDA path."""
    from pyaceqd_b200.pt_builder import build_qd_phonon_pt
    tls = tls_problem()
    pt = build_qd_phonon_pt(coupling_diag=tls.meta["coupling_diag"], dt=0.1, t_mem=6.4, a_e=5.0, temperature=4.0,
                            threshold=1e-8, backend="host")
    assert 16 <= pt.chi_max <= 48
    p = ChirpedPulse(tau_0=3.0, e_start=0.0, alpha=0, t0=10.0, e0=3.0)
    jobs = [Job(0.0, 25.0, 0.1, tables=make_tables([ChirpedPulse(tau_0=3.0, e_start=d, alpha=0, t0=10.0, e0=a)],
                                                    0.0, 25.0, 0.1))
            for a in (1.0, 3.0, 7.0) for d in (-1.0, 0.0, 1.5)]
    _check(engine, tls, pt, jobs, range(len(jobs)))
    bx = biexciton_problem(outputs=["|0><0|_4", "|1><1|_4", "|3><3|_4", "|0><3|_4"])
    ptb = build_qd_phonon_pt(coupling_diag=bx.meta["coupling_diag"], dt=0.5, t_mem=20.48, a_e=5.0, temperature=4.0,
                             threshold=1e-8, backend="host")
    assert ptb.chi_max >= 64
    pb = ChirpedPulse(tau_0=3.0, e_start=-2.0, alpha=0, t0=10.0, e0=6.0, polar_x=1.0)
    tabs = make_tables([pb], 0.0, 30.0, 0.5)
    jb = [Job(0.0, 30.0, 0.5, tables=tabs)]
    for t1 in (8.0, 12.5):
        jb.append(Job(0.0, t1 + 10.0, 0.5, tables=tabs, mtos=bx.parse_mtos(
            [{"operator": "|3><1|_4", "applyFrom": "_right", "time": t1},
             {"operator": "|1><3|_4", "applyFrom": "_left", "time": t1}])))
    _check(engine, bx, ptb, jb, range(len(jb)), min_signal=1e-3)
    _check(engine, bx, ptb, jb, range(len(jb)), min_signal=1e-3, kernel="splitk")


@pytest.mark.parametrize("which,name", [("fivels", "k_opbuild_dmma_cta<4>"), ("sixls", "k_opbuild_dmma_cta<5>")])
def test_cta_per_entry_operator_builder(engine, monkeypatch, which, name):
    """Pulse-area sweeps of the five- and six-level models (per-trajectory drives: every row needs its own two
    exponentials) through the tensor-core builder with one CTA per entry, against the oracle and against the group
    kernel (``ACEQD_OPBUILD_GROUP=1``); multi-time operators before and after a row included."""
    prob = fivels_problem() if which == "fivels" else sixls_problem()
    d = 5 if which == "fivels" else 6
    pt = synthetic_pt(24, len(prob.cls_keys), kind="unitary", scale=0.999)
    dt = 0.1
    jobs = []
    for k, a in enumerate(np.linspace(0.5, 9.0, 7)):
        p = ChirpedPulse(tau_0=0.6, e_start=-2.0, alpha=0.1 * k, t0=1.0, e0=a, polar_x=0.8)
        mt = []
        if k % 2:
            mt = prob.parse_mtos([{"operator": "|0><1|_%d" % d, "applyFrom": "_left", "time": 0.7, "applyBefore": "true"},
                                  {"operator": "|1><0|_%d" % d, "applyFrom": "_right", "time": 1.3}])
        jobs.append(Job(0.0, 2.5, dt, tables=make_tables([p], 0.0, 2.5, dt), mtos=mt))
    monkeypatch.delenv("ACEQD_OPBUILD_GROUP", raising=False)
    got = engine.run_jobs(prob, pt, jobs, kernel="dmma")
    assert engine.last_kernels()["opbuild"] == name, engine.last_kernels()
    monkeypatch.setenv("ACEQD_OPBUILD_GROUP", "1")
    grp = engine.run_jobs(prob, pt, jobs, kernel="dmma")
    assert engine.last_kernels()["opbuild"].startswith("k_opbuild<"), engine.last_kernels()
    monkeypatch.delenv("ACEQD_OPBUILD_GROUP", raising=False)
    for g, h, jb in zip(got, grp, jobs):
        assert np.abs(g - h).max() < 1e-12
        assert np.abs(g - oracle.propagate(prob, pt, jb)).max() < 1e-10
