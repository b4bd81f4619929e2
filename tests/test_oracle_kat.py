"""Known-answer tests that pin the oracle: the reference has no numeric
golden vectors for this path, so the restatement is checked against closed-form physics."""
import numpy as np
import pytest

import oracle
from helpers import make_tables, tls_problem
from pyaceqd_b200.jobs import FieldTable, Job
from pyaceqd_b200.opparser import parse_operator
from pyaceqd_b200.problem import MTO, build_problem
from pyaceqd_b200.process_tensor import trivial_pt
from pyaceqd_b200.pulses import ChirpedPulse, Pulse


def test_k1_lindblad_decay_exact():
    """x(t) = exp(-gamma t): constant L, so the Trotter halves are exact."""
    g = 0.037
    prob = build_problem(initial="|1><1|_2", lindblad_ops=[["|0><1|_2", g]], interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=["|0><0|_2", "|1><1|_2"])
    job = Job(0.0, 50.0, 0.1)
    out = oracle.propagate(prob, trivial_pt(1), job)
    t = job.times()
    assert out.shape == (2, 501)
    assert np.abs(out[1] - np.exp(-g * t)).max() < 1e-12
    assert np.abs(out[0] + out[1] - 1).max() < 1e-12


def test_k2_resonant_pi_pulse_and_area_law():
    """Resonant Gaussian pulse of area pi*e0: x_final = sin^2(pi e0 / 2) up to O(dt^2)."""
    prob = tls_problem(lindblad=False, phonons=False)
    for e0, tol in ((1.0, 1e-6), (0.5, 1e-6), (2.0, 1e-5), (1.37, 1e-5)):
        p = ChirpedPulse(tau_0=3, e_start=0, alpha=0, t0=15, e0=e0)
        job = Job(0.0, 30.0, 0.1, tables=make_tables([p], 0.0, 30.0, 0.1))
        out = oracle.propagate(prob, trivial_pt(1), job)
        assert abs(out[1, -1].real - np.sin(np.pi * e0 / 2) ** 2) < tol, e0


def test_k2_convergence_with_dt():
    prob = tls_problem(lindblad=False, phonons=False)
    errs = []
    for dt in (0.4, 0.2, 0.1):
        p = Pulse(tau=2.0, e_start=0.8, t0=10.0, e0=1.0)     # detuned: time ordering matters
        job = Job(0.0, 20.0, dt, tables=make_tables([p], 0.0, 20.0, dt, quantise=False))
        errs.append(oracle.propagate(prob, trivial_pt(1), job)[1, -1].real)
    ref_job = Job(0.0, 20.0, 0.0125, tables=make_tables([Pulse(tau=2.0, e_start=0.8, t0=10.0, e0=1.0)], 0.0, 20.0, 0.0125, quantise=False))
    ref = oracle.propagate(prob, trivial_pt(1), ref_job)[1, -1].real
    e = [abs(x - ref) for x in errs]
    assert e[0] > e[1] > e[2] and e[2] < 2e-4      # second-order-ish convergence


def test_k3_trace_and_hermiticity():
    prob = tls_problem(lindblad=True, dephasing=0.02, e_x=0.4, phonons=False)
    p = ChirpedPulse(tau_0=2, e_start=0.3, alpha=5, t0=8, e0=2.2)
    job = Job(0.0, 20.0, 0.1, tables=make_tables([p], 0.0, 20.0, 0.1))
    out = oracle.propagate(prob, trivial_pt(1), job)
    assert np.abs(out[0] + out[1] - 1).max() < 1e-12
    assert np.abs(out[2] - np.conj(out[3])).max() < 1e-12      # <|0><1|> = conj <|1><0|>
    assert out[1].real.min() > -1e-12 and out[1].real.max() < 1 + 1e-12


def test_output_rows_and_tail_indexing():
    """N = round((te - ta)/dt) steps -> N+1 rows; consumers slice from the end."""
    prob = tls_problem(phonons=False)
    for te, n in ((5.0, 50), (4.96, 50), (0.0, 0), (0.1, 1)):
        job = Job(0.0, te, 0.1)
        assert oracle.propagate(prob, trivial_pt(1), job).shape == (4, n + 1)
    job = Job(-3.0, 2.0, 0.25)
    assert np.allclose(job.times(), np.arange(-3.0, 2.0 + 1e-9, 0.25))


def test_k5_qrt_mto_equals_dynamical_map_powers():
    """Without phonons the MTO trajectory equals the quantum-regression result built from the
    dynamical map: <A(t1) B(t1+tau) C(t1)> = Tr[B E(tau) (C rho(t1) A)]  (CW drive => E(tau) = E(dt)^k)."""
    from pyaceqd_b200.pulses import CWLaser
    prob = build_problem(initial="|0><0|_2", lindblad_ops=[["|0><1|_2", 0.05]], interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=["|1><1|_2", "(|1><0|_2*|1><1|_2*|0><1|_2)", "|0><1|_2"])
    dt, t1, tau_max = 0.1, 3.0, 4.0
    cw = CWLaser(e0=0.12, e_start=0.0)
    tabs = make_tables([cw], 0.0, t1 + tau_max, dt, quantise=False)
    a, c = parse_operator("|1><0|_2", 2), parse_operator("|0><1|_2", 2)
    mtos = [MTO(prob.mto_superop(a, "_right"), t1, False), MTO(prob.mto_superop(c, "_left"), t1, False)]
    g = oracle.propagate(prob, trivial_pt(1), Job(0.0, t1 + tau_max, dt, tables=tabs, mtos=mtos))
    E = oracle.dynamical_map(prob, trivial_pt(1), Job(0.0, t1 + tau_max, dt, tables=tabs))
    k1 = int(round(t1 / dt))
    rho_t1 = E[k1] @ prob.rho0
    start = (c @ rho_t1.reshape(2, 2) @ a).reshape(-1)
    step = E[1]          # time-independent generator: one-step map
    v = start.copy()
    for k in range(1, int(round(tau_max / dt)) + 1):
        v = step @ v
        assert abs(prob.out_w[0] @ v - g[0, k1 + k]) < 1e-12
    # tau = 0 element comes from the product operator BEFORE the MTO acts
    assert abs(g[1, k1] - np.trace(a @ parse_operator("|1><1|_2", 2) @ c @ rho_t1.reshape(2, 2))) < 1e-13


def test_mto_semantics_left_right_sandwich_before():
    prob = build_problem(initial="|0><0|_2", interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2"])
    sx = parse_operator("|1><0|_2 + |0><1|_2", 2)
    job = Job(0.0, 0.3, 0.1, mtos=[MTO(prob.mto_superop(sx, ""), 0.1, False)])
    out = oracle.propagate(prob, trivial_pt(1), job)
    assert abs(out[0, 1] - 1) < 1e-14 and abs(out[1, 2] - 1) < 1e-14      # visible at time + dt
    job = Job(0.0, 0.3, 0.1, mtos=[MTO(prob.mto_superop(sx, ""), 0.1, True)])
    assert abs(oracle.propagate(prob, trivial_pt(1), job)[1, 1] - 1) < 1e-14   # applyBefore: visible at time
    up = parse_operator("|1><0|_2", 2)
    job = Job(0.0, 0.2, 0.1, mtos=[MTO(prob.mto_superop(up, "_left"), 0.0, False)])
    out = oracle.propagate(prob, trivial_pt(1), job)     # rho -> |1><0| rho = |1><0|  => rho_10 = 1
    assert abs(out[2, 1] - 1) < 1e-14 and abs(out[3, 1]) < 1e-14           # <|0><1|> = rho_10
    job = Job(0.0, 0.2, 0.1, mtos=[MTO(prob.mto_superop(up.conj().T, "_right"), 0.0, False)])
    out = oracle.propagate(prob, trivial_pt(1), job)     # rho -> rho |0><1| => rho_01 = 1
    assert abs(out[3, 1] - 1) < 1e-14 and abs(out[2, 1]) < 1e-14
    with pytest.raises(ValueError):
        oracle.propagate(prob, trivial_pt(1), Job(0.0, 0.2, 0.1, mtos=[MTO(np.eye(4), 5.0, False)]))


def test_field_sampling_rule():
    tab = FieldTable(1.0, 0.5, np.array([1.0, 3.0, 2.0 + 2j]))
    assert oracle.sample_field(tab, 0.0) == 1.0          # held before the table
    assert oracle.sample_field(tab, 1.25) == 2.0         # linear in between
    assert oracle.sample_field(tab, 1.75) == 2.5 + 1j
    assert oracle.sample_field(tab, 9.0) == 2.0 + 2j     # held after the table
    assert oracle.half_step_times(2.0, 0.1, "half_mid") == (2.025, 2.075)
