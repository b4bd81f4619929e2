"""Time-local-map chains: the GPU stand-ins for the reference's Fortran modules
against the literal NumPy restatement of the Fortran (oracle/tlmap_oracle.py).

CPU tests run the program builders on a NumPy interpreter of chain programs (host logic: index
schedules, run-length encoding, column-major conventions); GPU tests run the same cases on the
CUDA chain kernel through the C ABI.  Tolerance 1e-12 (identical operation order, FP64)."""
import numpy as np
import pytest

import tlmap_oracle as fo
from oracle_backend import oracle_backend, run_programs_numpy

TOL = 1e-12


def _maps(rng, NL, n, scale=0.9):
    m = rng.standard_normal((NL, NL, n)) + 1j * rng.standard_normal((NL, NL, n))
    for k in range(n):
        m[:, :, k] *= scale / np.linalg.norm(m[:, :, k], 2)
    return np.asfortranarray(m)


def _ops(rng, dim, k):
    return [rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)) for _ in range(k)]


def case_propagate_tau(dim):
    from pyaceqd_b200.two_time import propagate_tau_module as gm
    rng = np.random.default_rng(dim)
    NL = dim * dim
    dm = _maps(rng, NL, 40)
    rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    return gm.propagate_tau(dm, rho, 25, dim, 7), fo.propagate_tau(dm, rho, 25, dim, 7)


def case_onetime(dim):
    from pyaceqd_b200.two_time import propagate_tau_module as gm
    rng = np.random.default_rng(10 + dim)
    NL = dim * dim
    n_full, n_tau = 60, 20
    dm = _maps(rng, NL, n_full - 1)
    rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    A, B, C = _ops(rng, dim, 3)
    time = np.round(0.1 * np.arange(n_full), 6)
    sparse = time[[0, 1, 2, 5, 9, 14, 30, 38]]
    args = (dm, rho, n_tau, dim, A, B, C, time, sparse)
    return gm.calc_onetime_parallel(*args), fo.calc_onetime(*args)


def case_block(dim):
    from pyaceqd_b200.two_time import propagate_tau_module as gm
    rng = np.random.default_rng(20 + dim)
    NL = dim * dim
    n_tb, n_map, nx = 12, 7, 3
    block, dm_s = _maps(rng, NL, n_map), _maps(rng, NL, 1)[:, :, 0]
    rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    A, B, C = _ops(rng, dim, 3)
    time = np.round(0.1 * np.arange(40), 6)
    sparse = time[[0, 3, 6, 7, 11, 12, 13, 20]]          # before, at and beyond n_map / n_tb
    args = (block, dm_s, rho, n_tb, nx, dim, A, B, C, time, sparse)
    return gm.calc_onetime_parallel_block(*args), fo.calc_onetime_parallel_block(*args)


def case_phonon_block(dim):
    from pyaceqd_b200.two_time import propagate_tau_module as gm
    rng = np.random.default_rng(30 + dim)
    NL = dim * dim
    n_tb, n_map, nx, n_tauc = 12, 7, 3, 3
    sep1, sep2, dm_s = _maps(rng, NL, n_map), _maps(rng, NL, n_map), _maps(rng, NL, 1)[:, :, 0]
    taucs = np.asfortranarray(np.stack([_maps(rng, NL, n_map) for _ in range(n_tauc)], axis=2))
    rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    A, B, C = _ops(rng, dim, 3)
    time = np.round(0.1 * np.arange(40), 6)
    sparse = time[[0, 2, 5, 8, 10, 11, 12, 15]]
    args = (taucs, sep1, sep2, dm_s, rho, n_tb, nx, dim, A, B, C, time, sparse)
    return gm.calc_twotime_phonon_block(*args), fo.calc_twotime_phonon_block(*args)


def _timebin_inputs(dim, seed):
    rng = np.random.default_rng(seed)
    NL = dim * dim
    n_map, dt, tb = 10, 0.5, 16.0
    dm1, dm2 = _maps(rng, NL, n_map), _maps(rng, NL, n_map)
    E = _maps(rng, NL, 1, scale=0.97)[:, :, 0]
    pre = np.zeros((NL, NL, 7), dtype=complex, order="F")
    pre[:, :, 0] = E
    for i in range(1, 7):
        pre[:, :, i] = pre[:, :, i - 1] @ pre[:, :, i - 1]
    rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    t1 = np.array([0.0, 0.5, 1.0, 2.5, 4.0, 7.5, 12.0, 16.0])      # crosses the explicit-map range (5 ps)
    return dm1, dm2, rho, t1, pre, dt, tb, rng


def case_four_time(dim):
    from pyaceqd_b200.timebin import timebin_tl as gm
    dm1, dm2, rho, t1, pre, dt, tb, rng = _timebin_inputs(dim, 40 + dim)
    o = _ops(rng, dim, 4)
    args = (dm1, dm2, rho, t1, pre, dt, dim, *o, tb)
    return gm.four_time(*args), fo.four_time(*args)


def case_four_time_8op(dim, early=False, late=False):
    from pyaceqd_b200.timebin import timebin_tl as gm
    dm1, dm2, rho, t1, pre, dt, tb, rng = _timebin_inputs(dim, 50 + dim)
    o = _ops(rng, dim, 8)
    args = (dm1, dm2, rho, t1, pre, dt, dim, *o, early, late, tb)
    return gm.four_time_8op(*args), fo.four_time_8op(*args)


def case_dynamics(dim):
    from pyaceqd_b200.timebin import timebin_tl as gm
    dm1, dm2, rho, t1, pre, dt, tb, rng = _timebin_inputs(dim, 60 + dim)
    o = _ops(rng, dim, 3)
    a = gm.dynamics_t1(dm1, dm2, rho, t1, pre, dt, dim, tb), fo.dynamics_t1(dm1, dm2, rho, t1, pre, dt, dim, tb)
    b = (gm.dynamics_t1_t2(dm1, dm2, 1.0, 4.0, rho, t1, pre, dt, dim, tb, *o),
         fo.dynamics_t1_t2(dm1, dm2, 1.0, 4.0, rho, t1, pre, dt, dim, tb, *o))
    pt = (gm.utils.propagate_tb(1.0, 14.5, dt, rho, dm1, pre), fo.propagate_tb(1.0, 14.5, dt, rho, dm1, pre))
    return (np.concatenate([a[0].ravel(), b[0].ravel(), pt[0]]), np.concatenate([a[1].ravel(), b[1].ravel(), pt[1]]))


CASES = [("propagate_tau", case_propagate_tau), ("onetime", case_onetime), ("block", case_block),
         ("phonon_block", case_phonon_block), ("four_time", case_four_time), ("four_time_8op", case_four_time_8op),
         ("dynamics", case_dynamics)]


@pytest.mark.parametrize("name,fn", CASES)
@pytest.mark.parametrize("dim", [2, 4])
def test_program_builders_match_fortran_restatement(name, fn, dim):
    with oracle_backend():
        got, want = fn(dim)
    assert got.shape == want.shape and np.abs(got - want).max() < TOL


def test_8op_truncations_and_rle():
    from pyaceqd_b200.tlmap import rle
    with oracle_backend():
        for kw in ({"early": True}, {"late": True}):
            got, want = case_four_time_8op(2, **kw)
            assert np.abs(got - want).max() < TOL
    assert rle(np.array([3, 4, 5, 9, 9, 9, 2, 7, 8]), True) == [(3, 3, 1, 1), (9, 3, 1, 0), (2, 1, 1, 1), (7, 2, 1, 1)]
    assert rle(np.array([5]), False) == [(5, 1, 0, 1)] and rle(np.array([], dtype=int), True) == []


def test_quantum_regression_consistency():
    """Physics anchor of the restatement: for Markovian maps E = exp(L dt) the chain result equals
    Tr(B e^{L tau}[C rho(t) A]) computed directly (column-major bookkeeping included)."""
    import scipy.linalg
    rng = np.random.default_rng(1)
    dim, dt, n_full, n_tau = 2, 0.1, 30, 12
    H = np.array([[0.0, 0.4], [0.4, 0.3]])
    a = np.array([[0, 1], [0, 0]], dtype=complex)
    I = np.eye(dim)
    # column-major superoperators: vec(X M) = kron(I, X) vec(M), vec(M X) = kron(X^T, I) vec(M)
    Lsup = -1j * (np.kron(I, H) - np.kron(H.T, I)) + 0.2 * (np.kron(a.conj(), a) - 0.5 * np.kron(I, a.conj().T @ a)
                                                          - 0.5 * np.kron((a.conj().T @ a).T, I))
    E = scipy.linalg.expm(Lsup * dt)
    dm = np.asfortranarray(np.repeat(E[:, :, None], n_full - 1, axis=2))
    rho0 = np.array([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, 0.7]])
    A, B, C = _ops(rng, dim, 3)
    time = np.round(dt * np.arange(n_full), 6)
    sparse = time[[0, 4, 9]]
    res = fo.calc_onetime(dm, rho0.reshape(-1, order="F"), n_tau, dim, A, B, C, time, sparse)
    for i, ts in enumerate(sparse):
        rho_t = (scipy.linalg.expm(Lsup * ts) @ rho0.reshape(-1, order="F")).reshape(dim, dim, order="F")
        assert abs(res[i, 0] - np.trace(A @ B @ C @ rho_t)) < 1e-12
        for k in (1, 5, 12):
            r = (scipy.linalg.expm(Lsup * k * dt) @ (C @ rho_t @ A).reshape(-1, order="F")).reshape(dim, dim, order="F")
            assert abs(res[i, k] - np.trace(B @ r)) < 1e-12


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,fn", CASES)
@pytest.mark.parametrize("dim", [2, 4, 5, 6])
def test_chain_kernel_matches_fortran_restatement(name, fn, dim):
    got, want = fn(dim)
    assert got.shape == want.shape and np.abs(got - want).max() < TOL


@pytest.mark.gpu
def test_chain_kernel_generic_programs(engine):
    """Random programs (strides 0/1, several emitting segments, ragged chains, NL up to 64) against
    the NumPy interpreter."""
    from pyaceqd_b200.engine import TLSEG_DT
    rng = np.random.default_rng(9)
    for NL in (4, 25, 33, 64):
        mats = np.moveaxis(_maps(rng, NL, 50, scale=0.95), 2, 0)
        n_chains = 37
        v0 = rng.standard_normal((n_chains, NL)) + 1j * rng.standard_normal((n_chains, NL))
        seg_list, seg_off = [], [0]
        for c in range(n_chains):
            for _ in range(rng.integers(0, 5)):
                stride = int(rng.integers(0, 2))
                count = int(rng.integers(1, 12))
                start = int(rng.integers(0, 50 - count))
                seg_list.append((start, count, int(rng.integers(0, 2)), stride))
            seg_off.append(len(seg_list))
        segs = np.zeros(len(seg_list), dtype=TLSEG_DT)
        arr = np.asarray(seg_list, dtype=np.int32)
        segs["start"], segs["count"], segs["emit"], segs["stride"] = arr.T
        w = rng.standard_normal((3, NL)) + 1j * rng.standard_normal((3, NL))
        seg_off = np.asarray(seg_off, dtype=np.int64)
        n_emit = 30
        got_o, got_f = engine.tlmap_run(mats, v0, seg_off, segs, w=w, n_emit_max=n_emit, want_final=True)
        want_o, want_f = run_programs_numpy(mats, v0, seg_off, segs, w=w, n_emit_max=n_emit, want_final=True)
        assert np.abs(got_o - want_o).max() < TOL and np.abs(got_f - want_f).max() < TOL
