"""On-device process-tensor builder (``csrc/ptbuild.cu``): the C ABI loads and refuses to run
without a GPU (CPU tests); on a GPU its tensors reproduce the NumPy builder's physics -- same bond dimension, same
propagated observables, the brute-force influence functional and the independent-boson known answer (K4)."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

import oracle
from pyaceqd_b200 import constants
from pyaceqd_b200 import pt_builder as pb
from pyaceqd_b200 import pt_device
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import build_problem, coupling_classes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ptbuild_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "aceqd_ptbuild.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(aceqd_ptbuild_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 4
    lib = pt_device.load_library()
    for n in names:
        assert hasattr(lib, n), n


def test_device_builder_refuses_without_a_gpu():
    if pt_device.available():
        pytest.skip("a GPU is present")
    assert pb.resolve_backend(None) == "host" and pb.resolve_backend("device") == "device"
    with pytest.raises(pt_device.PtBuildError):
        pb.build_qd_phonon_pt([0.0, 1.0], dt=0.1, t_mem=1.0, backend="device")
    with pytest.raises(ValueError):
        pb.resolve_backend("fpga")


def _spectral():
    w = np.linspace(0.0, 7.0 / constants.hbar, 20001)
    return w, pb.qd_phonon_spectral_density(w, 5.0)


@pytest.mark.gpu
def test_eta_coefficients_on_the_device():
    w, J = _spectral()
    for T in (4.0, 77.0, 0.0):
        host = pb.eta_coefficients(J, w, 0.1, 64, T)
        dev = pt_device.eta_coefficients(J, w, 0.1, 64, T)
        assert np.abs(dev - host).max() < 1e-13 * np.abs(host).max(), T


@pytest.mark.gpu
@pytest.mark.parametrize("svd", [0, 1])
def test_device_mps_equals_bruteforce_influence_functional(svd):
    """Same check as tests/test_pt_builder.py holds for the NumPy builder: the uniform MPS reproduces the discretised
    influence functional of every path (exact at threshold 1e-14)."""
    _, keys = coupling_classes(np.array([0.0, 1.0, 2.0]))
    eta = np.array([0.2 + 0.05j, 0.1 - 0.04j, 0.03 + 0.01j])     # memory 2: bond dimension 81 at this threshold
    K = len(eta) - 1
    os.environ["ACEQD_PT_SVD"] = str(svd)
    try:
        pt = pb.uniform_pt(keys, eta, dt=0.5, threshold=1e-14, shift_rate=0.3, backend="device")
    finally:
        del os.environ["ACEQD_PT_SVD"]
    assert pt.meta["device_build"]["svd"] == ("gesvd", "gesvdp")[svd]
    A, q = pt.slices[0], pt.closures[0]
    I, i0 = pb.influence_factors(keys, eta, 0.5, 0.3)
    rng = np.random.default_rng(1)
    for _ in range(300):
        cs = rng.integers(0, 9, size=6)
        F = np.prod([i0[c] for c in cs])
        for n in range(6):
            for k in range(1, K + 1):
                if n - k >= 0:
                    F *= I[k][cs[n], cs[n - k]]
        v = np.zeros(A.shape[1], complex)
        v[0] = 1
        for c in cs:
            v = v @ A[c]
        # the polar-decomposition SVD resolves singular values to ~1e-8 of the largest one only
        assert abs(v @ q - F) < (1e-10, 1e-6)[svd]


@pytest.mark.gpu
def test_device_pt_matches_the_numpy_builder_through_the_cuda_path(engine):
    """TLS (threshold 1e-8) and biexciton-class (9 classes, threshold 1e-7) PTs built on the device and on the host:
    same bond dimension, and the observables of driven runs propagated by the step kernel agree far below the
    truncation level the threshold itself allows."""
    from helpers import biexciton_problem, make_tables, tls_problem
    from pyaceqd_b200.pulses import ChirpedPulse
    cases = [(tls_problem(), dict(dt=0.1, t_mem=6.4, a_e=5.0, temperature=4.0, threshold=1e-8), 25.0,
              ChirpedPulse(tau_0=3.0, e_start=0.5, alpha=0, t0=10.0, e0=3.0)),
             (biexciton_problem(outputs=["|0><0|_4", "|1><1|_4", "|3><3|_4", "|0><3|_4"]),
              dict(dt=0.5, t_mem=10.0, a_e=5.0, temperature=4.0, threshold=1e-7), 30.0,
              ChirpedPulse(tau_0=3.0, e_start=-2.0, alpha=0, t0=10.0, e0=6.0, polar_x=1.0))]
    for prob, kw, tend, pulse in cases:
        host = pb.build_qd_phonon_pt(coupling_diag=prob.meta["coupling_diag"], backend="host", **kw)
        dev = pb.build_qd_phonon_pt(coupling_diag=prob.meta["coupling_diag"], backend="device", **kw)
        assert dev.meta["backend"] == "device" and dev.meta["device_build"]["svd_ms"] > 0
        assert abs(dev.chi_max - host.chi_max) <= 1, (dev.chi_max, host.chi_max)
        job = Job(0.0, tend, kw["dt"], tables=make_tables([pulse], 0.0, tend, kw["dt"]))
        a = engine.run_jobs(prob, host, [job])[0]
        b = engine.run_jobs(prob, dev, [job])[0]
        assert np.abs(a).max() > 0.1
        assert np.abs(a - b).max() < 1e-9, np.abs(a - b).max()
        assert np.abs(b - oracle.propagate(prob, dev, job)).max() < 1e-10


@pytest.mark.gpu
def test_k4_independent_boson_model_with_a_device_built_pt(engine):
    """K4 end to end on the GPU: device-built PT + step kernel against the analytic dephasing
    rho_10(t) = rho_10(0) exp(-Phi(t))."""
    temperature = 4.0
    pt = pb.build_qd_phonon_pt([0.0, 1.0], dt=0.1, t_mem=6.4, a_e=5.0, temperature=temperature, threshold=1e-8,
                               backend="device")
    prob = build_problem(boson_op="1*|1><1|_2", rho0=np.array([[0.5, 0.5], [0.5, 0.5]]), dim=2,
                         interaction_ops=[["|1><0|_2", "x"]], output_ops=["|0><1|_2", "|1><1|_2", "|0><0|_2"])
    job = Job(0.0, 15.0, 0.1)
    out = engine.run_jobs(prob, pt, [job])[0]
    tt = job.times()
    w, J = _spectral()
    coth = np.zeros_like(w)
    coth[1:] = 1 / np.tanh(constants.hbar * w[1:] / (2 * constants.kB * temperature))
    jw2 = np.zeros_like(w)
    jw2[1:] = J[1:] / w[1:] ** 2
    phi = np.array([np.trapezoid(jw2 * (coth * (1 - np.cos(w * x)) + 1j * np.sin(w * x)), w) for x in tt])
    assert np.abs(out[0] - 0.5 * np.exp(-phi)).max() < 5e-4
    assert np.abs(out[1] - 0.5).max() < 1e-5
