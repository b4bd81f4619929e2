"""Reference-facing API on the GPU: adapters (tls, biexciton, ...), BatchExecutor, two_time
workflows -- checked against the oracle driven with the same operator strings."""
import numpy as np
import pytest

import oracle
from pyaceqd_b200.batch import BatchExecutor, wait
from pyaceqd_b200.general_system import general_system as gs
from pyaceqd_b200.jobs import FieldTable, Job
from pyaceqd_b200.problem import build_problem
from pyaceqd_b200.process_tensor import synthetic_pt, trivial_pt
from pyaceqd_b200.pulses import ChirpedPulse

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _oracle_tls(t0, t1, pulse, dt, lindblad, gamma_e=0.01, mtos=None, outputs=None, pt=None):
    prob = build_problem(boson_op="1.000*|1><1|_2" if pt is not None else None, initial="|0><0|_2",
                         lindblad_ops=[["|0><1|_2", gamma_e]] if lindblad else [],
                         interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=outputs or ["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2"])
    t = np.arange(t0, t1, dt)
    px, _ = gs.sample_pulses(t, [pulse])
    job = Job(t0, t1, dt, tables={"x": FieldTable(t0, dt, px)}, mtos=prob.parse_mtos(mtos))
    return oracle.propagate(prob, pt or trivial_pt(len(prob.cls_keys)), job)


def test_tls_drop_in_return_layout():
    """cfg1 of SURVEY 8d (shortened): t, g, x, pgx, pxg = tls(...)."""
    from pyaceqd_b200.two_level_system.tls import tls
    p = ChirpedPulse(tau_0=3, e_start=0, alpha=0, t0=12, e0=1)
    res = tls(0, 40, p, dt=0.1, lindblad=True)
    t, g, x, pgx, pxg = res
    assert res.shape == (5, 401) and res.dtype == complex
    assert np.allclose(t.real, 0.1 * np.arange(401))
    ref = _oracle_tls(0.0, 40.0, p, 0.1, True)
    assert np.abs(res[1:] - ref).max() < TOL
    assert abs(x[-1].real - np.exp(-0.01 * 22)) < 0.05 and np.abs(g + x - 1).max() < 1e-12


def test_tls_phonons_with_pt_file(tmp_path):
    from pyaceqd_b200.two_level_system.tls import tls
    pt = synthetic_pt(16, 4, kind="unitary", scale=0.999)
    f = str(tmp_path / "synthetic.pt")
    pt.save(f)
    p = ChirpedPulse(tau_0=2, e_start=0.4, alpha=0, t0=6, e0=2.0)
    res = tls(0, 12, p, dt=0.1, phonons=True, pt_file=f, lindblad=True)
    ref = _oracle_tls(0.0, 12.0, p, 0.1, True, pt=pt)
    assert np.abs(res[1:] - ref).max() < TOL


def test_batch_executor_sweep_equals_eager_calls():
    """rabi_rotations-style fan-out (reference two_level_system/rabi_rotations.py:172-198)."""
    from pyaceqd_b200.two_level_system.tls import tls
    areas = np.linspace(0.5, 4.0, 9)
    futures = []
    with BatchExecutor(max_workers=15) as ex:
        for i, a in enumerate(areas):
            p1 = ChirpedPulse(tau_0=2.0, e_start=0.0, alpha=0, e0=a, polar_x=1.0, t0=8.0)
            futures.append(ex.submit(tls, 0, 16.0, p1, lindblad=False, suffix=i))
        wait(futures)
    final = np.array([f.result()[2][-1].real for f in futures])
    assert np.abs(final - np.sin(np.pi * areas / 2) ** 2).max() < 1e-3     # Rabi rotations (pulse tails cut at 4 tau)
    p1 = ChirpedPulse(tau_0=2.0, e_start=0.0, alpha=0, e0=areas[3], polar_x=1.0, t0=8.0)
    assert np.abs(tls(0, 16.0, p1, lindblad=False) - futures[3].result()).max() < 1e-13


def test_three_op_two_time_g2_grid():
    """G2(t, tau) of a driven TLS (reference two_time/correlations.py:227-270) vs per-t oracle runs."""
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.correlations import three_op_two_time
    p = ChirpedPulse(tau_0=1.5, e_start=0, alpha=0, t0=4, e0=3)
    t_axis = np.round(np.arange(0.0, 8.0, 0.5), 6)
    opts = {"lindblad": True, "phonons": False, "gamma_e": 0.05}
    t1, tau, G = three_op_two_time(tls, t_axis, p, tau_max=5.0, dt=0.1, options=dict(opts))
    assert G.shape == (len(t_axis), 51) and np.allclose(tau, np.linspace(0, 5.0, 51))
    outs = ["|1><1|_2", "(|1><0|_2*|1><1|_2*|0><1|_2)"]
    worst = 0.0
    for j, t1_j in enumerate(t_axis):
        mtos = [{"operator": "|1><0|_2", "applyFrom": "_right", "applyBefore": "false", "time": t1_j},
                {"operator": "|0><1|_2", "applyFrom": "_left", "applyBefore": "false", "time": t1_j}]
        # like the reference, every run sees its OWN pulse file: samples on np.arange(0, t_end, dt), end value held
        # (general_system.py:213) -- although the engine shares one table and one trunk across all t1
        prob = build_problem(initial="|0><0|_2", lindblad_ops=[["|0><1|_2", 0.05]],
                             interaction_ops=[["|1><0|_2", "x"]], output_ops=outs)
        tt = np.arange(0.0, float(t1_j + 5.0), 0.1)
        px, _ = gs.sample_pulses(tt, [p])
        job = Job(0.0, float(t1_j + 5.0), 0.1, tables={"x": FieldTable(0.0, 0.1, px)}, mtos=prob.parse_mtos(mtos))
        ref = oracle.propagate(prob, trivial_pt(1), job)
        worst = max(worst, np.abs(G[j, 1:] - ref[0][-50:]).max(), abs(G[j, 0] - ref[1][-51]))
    assert worst < TOL
    assert np.abs(G[:, 0].imag).max() < 1e-12 and G[:, 0].real.min() > -1e-12     # G2(t,0) = <n(n-1)>-like, real


def test_biexciton_sixls_darkmodel_adapters_run():
    from pyaceqd_b200.four_level_system.dark_model import darkmodel_new
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.six_level_system.linear import sixls_linear
    p = ChirpedPulse(tau_0=2.0, e_start=-2.0, alpha=0, t0=8, e0=4.0, polar_x=0.8)
    r4 = biexciton(0, 20, p, dt=0.25, lindblad=True, delta_b=4, delta_xy=0.05)
    assert r4.shape == (5, 81) and np.abs(r4[1:].sum(axis=0) - 1).max() < 1e-11
    assert r4[4][-1].real > 0.05      # two-photon resonant pulse populates the biexciton
    r5 = darkmodel_new(0, 10, p, dt=0.25, lindblad=True)
    assert r5.shape == (6, 41) and np.abs(r5[1:].sum(axis=0) - 1).max() < 1e-11
    t, rho = sixls_linear(0, 6, p, dt=0.25, lindblad=True, bx=2.0, output_dm=True)
    assert rho.shape == (25, 6, 6)
    assert np.abs(np.trace(rho, axis1=1, axis2=2) - 1).max() < 1e-11
    assert np.abs(rho - np.conj(np.transpose(rho, (0, 2, 1)))).max() < 1e-12
    # rotating frame: populations are frame independent up to the O(dt^2) error of sampling the
    # 2 meV carrier in the lab frame (0.088 at dt=0.25, 0.0035 at dt=0.05 -- same as the oracle)
    lab = biexciton(0, 20, p, dt=0.05, lindblad=True, delta_b=4, delta_xy=0.05)
    rot = biexciton(0, 20, p, dt=0.05, lindblad=True, delta_b=4, delta_xy=0.05, rf=True)
    assert np.abs(rot[1:5] - lab[1:5]).max() < 5e-3


def test_calc_dynmap_and_get_M_t():
    from scipy.linalg import expm
    from pyaceqd_b200.two_level_system.tls import tls
    p = ChirpedPulse(tau_0=1.5, e_start=0.2, alpha=0, t0=3, e0=1.5)
    res, E = tls(0, 6, p, dt=0.1, lindblad=True, calc_dynmap=True)
    assert E.shape == (60, 4, 4) and res.shape == (5, 61)     # E[i] = E_{t_{i+1}, t_0} (tools.py:470-479)
    rho_t = E @ np.array([1, 0, 0, 0], dtype=complex)        # map applied to |0><0|
    assert np.abs(rho_t[:, 3] - res[2][1:]).max() < 1e-12    # x population
    from pyaceqd_b200.tools import calc_tl_dynmap_pseudo
    tl = calc_tl_dynmap_pseudo(E, res[0].real)
    v = np.array([1, 0, 0, 0], dtype=complex)
    for k in range(len(tl)):
        v = tl[k] @ v
        assert abs(v[3] - res[2][k + 1]) < 1e-9              # time-local maps reproduce the trajectory
    M = tls(0, 6, p, dt=0.1, lindblad=True, get_M_t=1.0)
    assert M.shape == (4, 4) and np.abs(np.ones(4) @ np.eye(2).reshape(-1)[:, None] * 0).max() == 0
    tr = np.eye(2).reshape(-1)
    assert np.abs(tr @ M - tr).max() < 1e-12                 # trace preserving propagator


def test_fused_tail_reduction_equals_host_trapezoid(engine):
    """Workflow-level fusion: the tau integral of every run's tail is taken on the device
    (k_tail_reduce) where the step kernel left its outputs -- same numbers as np.trapz over the shipped rows."""
    from helpers import biexciton_problem, make_tables
    from pyaceqd_b200.engine import tail_trapezoid
    from pyaceqd_b200.process_tensor import synthetic_pt
    prob = biexciton_problem(outputs=["|1><1|_4", "|3><3|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)", "(|3><1|_4*|3><3|_4*|1><3|_4)"])
    pt = synthetic_pt(24, len(prob.cls_keys), kind="unitary", scale=0.999)
    dt, tend = 0.25, 9.0
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
    tabs = make_tables([p], 0.0, tend, dt)
    jobs = []
    for t1 in (0.0, 1.0, 2.5, 4.0, 8.75, 9.0):
        mt = prob.parse_mtos([{"operator": "|3><1|_4", "applyFrom": "_right", "time": t1},
                              {"operator": "|1><3|_4", "applyFrom": "_left", "time": t1}])
        jobs.append(Job(0.0, tend, dt, tables=tabs, mtos=mt, tail_rows=int(round((tend - t1) / dt)) + 1))
    pairs = [(0, 2), (1, 3)]
    full = engine.run_jobs(prob, pt, jobs)
    fused = engine.run_jobs(prob, pt, jobs, tail_reduce=(pairs, dt))
    assert engine.last_kernels()["other"] == "k_tail_reduce"
    for f, r, jb in zip(full, fused, jobs):
        want = tail_trapezoid(f, pairs, dt)
        assert r.shape == (2,) and np.abs(r - want).max() < 1e-12, jb.tail_rows
    assert np.abs(fused[-1]).max() == 0.0 and max(np.abs(r).max() for r in fused) > 1e-3    # one sample spans no interval
