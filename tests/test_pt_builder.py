"""PT builder: exactness of the uniform MPS against the brute-force influence
functional, and the independent-boson known answer end to end through the
oracle's slice contraction + closure."""
import itertools

import numpy as np
import pytest

import oracle
from pyaceqd_b200 import constants
from pyaceqd_b200 import pt_builder as pb
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import build_problem, coupling_classes
from pyaceqd_b200.process_tensor import ProcessTensor


def test_uniform_mps_equals_bruteforce_influence_functional():
    keys = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], float)
    rng = np.random.default_rng(0)
    K = 3
    eta = (rng.standard_normal(K + 1) * 0.3 + 1j * rng.standard_normal(K + 1) * 0.3) * np.array([1, 0.6, 0.3, 0.1])
    eta[0] = abs(eta[0].real) + 1j * eta[0].imag
    pt = pb.uniform_pt(keys, eta, dt=0.1, threshold=1e-14)
    A, q = pt.slices[0], pt.closures[0]
    I, i0 = pb.influence_factors(keys, eta, 0.1)
    N = 5
    worst = 0.0
    for cs in itertools.product(range(4), repeat=N):
        F = 1.0 + 0j
        for n in range(N):
            F *= i0[cs[n]]
            for k in range(1, K + 1):
                if n - k >= 0:
                    F *= I[k][cs[n], cs[n - k]]
        v = np.zeros(A.shape[1], complex)
        v[0] = 1
        for c in cs:
            v = v @ A[c]
        worst = max(worst, abs(v @ q - F))
    assert worst < 1e-12


def test_three_level_classes_bruteforce():
    """biexciton-type coupling {0,1,2}: 9 classes, memory 2."""
    _, keys = coupling_classes(np.array([0.0, 1.0, 2.0]))
    eta = np.array([0.2 + 0.05j, 0.1 - 0.04j, 0.03 + 0.01j])
    pt = pb.uniform_pt(keys, eta, dt=0.5, threshold=1e-14, shift_rate=0.3)
    A, q = pt.slices[0], pt.closures[0]
    I, i0 = pb.influence_factors(keys, eta, 0.5, 0.3)
    rng = np.random.default_rng(1)
    for _ in range(300):
        cs = rng.integers(0, 9, size=6)
        F = np.prod([i0[c] for c in cs])
        for n in range(6):
            for k in (1, 2):
                if n - k >= 0:
                    F *= I[k][cs[n], cs[n - k]]
        v = np.zeros(A.shape[1], complex)
        v[0] = 1
        for c in cs:
            v = v @ A[c]
        assert abs(v @ q - F) < 1e-11


def test_spectral_density_and_eta_sanity():
    w = np.linspace(0, 7.0 / constants.hbar, 20001)
    J = pb.qd_phonon_spectral_density(w, 5.0)
    assert J[0] == 0 and J.min() >= 0
    wmax = w[np.argmax(J)] * constants.hbar
    assert 0.8 < wmax < 2.5                       # LA-phonon sideband maximum around 1-2 meV for a_e = 5 nm
    assert J[-1] < 1e-6 * J.max()                 # negligible at the 7 meV cutoff (Boson_E_max)
    S = np.trapezoid(J[1:] / w[1:] ** 2, w[1:])   # Huang-Rhys factor
    assert 0.01 < S < 0.1
    eta = pb.eta_coefficients(J, w, 0.1, 64, 4.0)
    assert eta[0].real > 0 and abs(eta[64]) < 1e-3 * abs(eta[1])     # memory decays within 6.4 ps
    # smaller dots couple more strongly (tls.py docstring: "the smaller, the stronger")
    assert np.trapezoid(pb.qd_phonon_spectral_density(w, 3.0), w) > np.trapezoid(J, w)


@pytest.mark.parametrize("temperature,threshold,tol", [(4.0, 1e-8, 5e-4), (20.0, 1e-10, 1e-3)])
def test_k4_independent_boson_model(temperature, threshold, tol):
    """No driving: rho_10(t) = rho_10(0) exp(-Phi(t)),
    Phi = int dw J/w^2 [coth(hbar w/2kT)(1 - cos wt) + i sin wt]  (polaron shift subtracted)."""
    pt = pb.build_qd_phonon_pt([0.0, 1.0], dt=0.1, t_mem=6.4, a_e=5.0, temperature=temperature, threshold=threshold)
    assert 8 <= pt.chi_max <= 256
    prob = build_problem(boson_op="1*|1><1|_2", rho0=np.array([[0.5, 0.5], [0.5, 0.5]]), dim=2,
                         interaction_ops=[["|1><0|_2", "x"]], output_ops=["|0><1|_2", "|1><1|_2", "|0><0|_2"])
    job = Job(0.0, 15.0, 0.1)
    out = oracle.propagate(prob, pt, job)
    tt = job.times()
    w = np.linspace(0, 7.0 / constants.hbar, 20001)
    J = pb.qd_phonon_spectral_density(w, 5.0)
    coth = np.zeros_like(w)
    coth[1:] = 1 / np.tanh(constants.hbar * w[1:] / (2 * constants.kB * temperature))
    jw2 = np.zeros_like(w)
    jw2[1:] = J[1:] / w[1:] ** 2
    phi = np.array([np.trapezoid(jw2 * (coth * (1 - np.cos(w * x)) + 1j * np.sin(w * x)), w) for x in tt])
    assert np.abs(out[0] - 0.5 * np.exp(-phi)).max() < tol          # <|0><1|> = rho_10
    assert np.abs(out[1] - 0.5).max() < 1e-5 and np.abs(out[1] + out[2] - 1).max() < 1e-5
    assert abs(out[0][-1]) < abs(out[0][0])                          # phonon-induced initial dephasing


def test_pt_save_load_roundtrip(tmp_path):
    pt = pb.build_qd_phonon_pt([0.0, 1.0, 1.0, 2.0], dt=0.5, t_mem=4.0, threshold=1e-6)
    assert pt.n_cls == 9
    f = str(tmp_path / "b_linear_3.0nm_4k_th6_tmem4.0_dt0.5.ptr")
    pt.save(f)
    back = ProcessTensor.load(f)
    assert np.array_equal(back.slices[0], pt.slices[0]) and np.array_equal(back.closures[0], pt.closures[0])
    assert np.array_equal(back.keys, pt.keys) and back.n_initial == 0 and back.dt == 0.5
    # class lookup by coupling eigenvalue pair
    _, keys = coupling_classes(np.array([0.0, 1.0, 1.0, 2.0]))
    assert back.block_of_class(keys).tolist() == list(range(9))
    with pytest.raises(ValueError):
        back.block_of_class(np.array([[0.0, 3.0]]))


def test_spectral_density_file_round_trip(tmp_path):
    """Boson_J_print / Boson_J_from_file (reference general_system.py:178-179,186-187): a PT built from the printed
    QDPhonon spectral density equals the PT built from the analytic one (up to the table's interpolation error)."""
    from pyaceqd_b200.pt_builder import (build_pt_from_spectral_density_file, build_qd_phonon_pt, read_spectral_density,
                                         write_spectral_density)
    f = str(tmp_path / "J_omega.dat")
    write_spectral_density(f, a_e=5.0)
    w, J = read_spectral_density(f)
    assert len(w) == 2000 and abs(w[-1] * 0.6582119569 - 15.0) < 1e-9 and J[0] == 0.0 and J.max() > 0.05
    a = build_qd_phonon_pt([0.0, 1.0], dt=0.2, t_mem=4.0, a_e=5.0, temperature=4.0, threshold=1e-7)
    b = build_pt_from_spectral_density_file(f, [0.0, 1.0], dt=0.2, t_mem=4.0, temperature=4.0, threshold=1e-7, e_max=7.0)
    assert abs(a.chi_max - b.chi_max) <= 1
    # compare what the tensors DO: dephasing of the coherence class over 30 steps
    def decay(pt):
        k = pt.block_of_class(np.array([[1.0, 0.0]]))[0]
        v = np.zeros(pt.chi_max, dtype=complex)
        v[0] = 1.0
        out = []
        for _ in range(30):
            v = v @ pt.slices[0][k]
            out.append(v @ pt.closures[0])
        return np.array(out)
    assert np.abs(decay(a) - decay(b)).max() < 2e-5
