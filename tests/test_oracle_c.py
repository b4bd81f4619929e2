"""The C restatement (timed CPU baseline) against the NumPy oracle and scipy's expm."""
import numpy as np
from scipy.linalg import expm

import oracle
import oracle_c
from helpers import biexciton_problem, sweep_jobs, tls_problem
from pyaceqd_b200.process_tensor import synthetic_growing_pt, synthetic_pt, trivial_pt


def test_c_expm_matches_scipy():
    rng = np.random.default_rng(2)
    for n in (2, 4, 9, 16):
        a = (rng.standard_normal((4, n, n)) + 1j * rng.standard_normal((4, n, n))) * rng.uniform(0.01, 8, (4, 1, 1))
        got = oracle_c.expm(a)
        for i in range(4):
            ref = expm(a[i])
            assert np.abs(got[i] - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_c_propagation_matches_numpy_oracle():
    cases = [(tls_problem(), synthetic_pt(16, 4, n_slices=2, kind="unitary")),
             (tls_problem(phonons=False), trivial_pt(1)),
             (biexciton_problem(), synthetic_growing_pt(12, 9, 4, 3))]
    for prob, pt in cases:
        jobs = sweep_jobs(2, 3, t_end=3.0)
        got = oracle_c.propagate_sweep(prob, pt, jobs)
        for g, j in zip(got, jobs):
            assert np.abs(g - oracle.propagate(prob, pt, j)).max() < 1e-12


def test_c_threads_do_not_change_results():
    prob, pt = tls_problem(), synthetic_pt(8, 4, kind="unitary")
    jobs = sweep_jobs(3, 3, t_end=2.0)
    a = oracle_c.propagate_sweep(prob, pt, jobs, n_threads=1)
    b = oracle_c.propagate_sweep(prob, pt, jobs, n_threads=4)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
