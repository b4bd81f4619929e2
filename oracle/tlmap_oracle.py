"""CPU ORACLE (test infrastructure, NOT a product path) for the time-local-map chains.

Literal NumPy restatement of the reference's Fortran helpers, subroutine by subroutine, with
1-based loop bounds translated to 0-based and nothing else changed:

* ``pyaceqd/two_time/propagate_tau.f90``: ``propagate_tau`` :3-19, ``calc_onetime`` /
  ``calc_onetime_parallel`` :43-187, ``calc_onetime_parallel_block`` :189-295,
  ``calc_twotime_phonon_block`` :374-536;
* ``pyaceqd/timebin/timebin_tl.f90``: ``fast_propagate`` :23-47, ``propagate_tb`` :50-77,
  ``apply_left`` / ``apply_right`` :101-121, ``four_time`` :145-214, ``four_time_8op`` :216-303,
  ``dynamics_t1`` :305-342, ``dynamics_t1_t2`` :344-397.

The Fortran sources cannot be compiled here (no gfortran in the image), so parity for these
routines is this restatement vs the CUDA path; the restatement is additionally checked against
plain physics (quantum-regression consistency with full trajectories) in the tests.
Array layouts are the ones f2py callers pass: maps as ``[NL, NL, n]`` (Fortran order), vectors
``[NL]`` that Fortran reshapes COLUMN-major to ``[dim, dim]``.
Only ``tests/`` may import this module.
"""
import numpy as np


def _mat(v, dim):
    return np.asarray(v).reshape(dim, dim, order="F")


def _vec(m):
    return np.asarray(m).reshape(-1, order="F")


def _tr(m):
    return np.trace(m)


def propagate_tau(dm_tl, rho_init, n_tau, dim, j_start):
    """propagate_tau.f90:3-19 -- rho_out[:, k+1] = dm_tl[:, :, j_start + k] rho_out[:, k]  (k 1-based:
    dm_tl(:,:,j_start+k), i.e. 0-based index j_start + k - 1 for k = 1..n_tau)."""
    out = np.zeros((dim * dim, n_tau + 1), dtype=complex)
    out[:, 0] = rho_init
    for k in range(1, n_tau + 1):
        out[:, k] = dm_tl[:, :, j_start + k - 1] @ out[:, k - 1]
    return out


def calc_onetime(dm_tl, rho_init, n_tau, dim, opA, opB, opC, time, time_sparse):
    """propagate_tau.f90:43-108 (and its OpenMP twin :110-187): result[i, 0] = Tr(A B C rho(t_i)),
    result[i, k] = Tr(B E..E (C rho A)) along tau."""
    n_t = len(time_sparse)
    res = np.zeros((n_t, n_tau + 1), dtype=complex)
    v = np.array(rho_init, dtype=complex)
    j = 0
    for i in range(n_t):
        while time[j] < time_sparse[i]:
            v = dm_tl[:, :, j] @ v
            j += 1
        m = _mat(v, dim)
        res[i, 0] = _tr(opA @ (opB @ (opC @ m)))
        r = _vec((opC @ m) @ opA)
        for k in range(2, n_tau + 2):                # Fortran k = 2 .. n_tau+1, map index j-2+k (1-based)
            r = dm_tl[:, :, (j + 1) - 2 + k - 1] @ r
            res[i, k - 1] = _tr(opB @ _mat(r, dim))
    return res


calc_onetime_parallel = calc_onetime


def calc_onetime_parallel_block(dm_block, dm_s, rho_init, n_tb, nx_tau, dim, opa, opb, opc, time, time_sparse):
    """propagate_tau.f90:189-295 -- maps are periodic with period n_tb: the first n_map of a period
    are dm_block, the rest the stationary map dm_s."""
    n_map = dm_block.shape[2]
    n_t = len(time_sparse)
    K = nx_tau * n_tb
    res = np.zeros((n_t, K + 1), dtype=complex)
    v = np.array(rho_init, dtype=complex)
    j = 1                                             # kept 1-based like the Fortran
    for i in range(n_t):
        while time[j - 1] < time_sparse[i]:
            v = (dm_block[:, :, j - 1] if j <= n_map else dm_s) @ v
            j += 1
        m = _mat(v, dim)
        res[i, 0] = _tr(opa @ (opb @ (opc @ m)))
        r = _vec((opc @ m) @ opa)
        jj = j
        for k in range(2, K + 2):
            r = (dm_block[:, :, jj - 1] if jj <= n_map else dm_s) @ r
            res[i, k - 1] = _tr(opb @ _mat(r, dim))
            jj += 1
            if jj == n_tb + 1:
                jj = 1
    return res


def calc_twotime_phonon_block(dm_taucs2, dm_sep1, dm_sep2, dm_s, rho_init, n_tb, nx_tau, dim, opa, opb, opc,
                              time, time_sparse):
    """propagate_tau.f90:374-536 -- the first period after the operator time uses chain-specific maps
    (dm_taucs2[:, :, i, :] for the first n_tauc chains, dm_sep2 for the others), later periods dm_sep1;
    no operator is applied to the branch start (:455-459) and the trace uses transpose(opB) (:484)."""
    n_map = dm_sep1.shape[2]
    n_tauc = dm_taucs2.shape[2]
    n_t = len(time_sparse)
    K = nx_tau * n_tb
    res = np.zeros((n_t, K + 1), dtype=complex)
    v = np.array(rho_init, dtype=complex)
    j = 1
    buf, j_arr = [], []
    for i in range(n_t):
        while time[j - 1] < time_sparse[i]:
            v = (dm_sep1[:, :, j - 1] if j <= n_map else dm_s) @ v
            j += 1
        res[i, 0] = _tr(opa @ (opb @ (opc @ _mat(v, dim))))
        buf.append(v.copy())
        j_arr.append(j)
    for i in range(n_t):
        r = buf[i]
        jj, j_start, use2 = 1, j_arr[i], True
        for k in range(2, K + 2):
            if jj <= n_map:
                if use2:
                    mp = dm_taucs2[:, :, i, jj - 1] if i < n_tauc else dm_sep2[:, :, jj - 1]
                else:
                    mp = dm_sep1[:, :, jj - 1]
            else:
                mp = dm_s
            r = mp @ r
            res[i, k - 1] = _tr(opb.T @ _mat(r, dim))
            jj += 1
            if jj + j_start == n_tb + 1:
                j_start, jj, use2 = 0, 1, False
    return res


# ------------------------------------------------------------------------------------ timebin_tl.f90
def _round6(x):
    return np.rint(x * 1_000_000) / 1_000_000


def fast_propagate(rho, dm_tl_precalc, n_steps):
    """timebin_tl.f90:23-47 -- binary decomposition of n_steps over precomputed powers E^(2^i)."""
    out = np.array(rho, dtype=complex)
    i, n = 0, int(n_steps)
    while n > 0:
        if n & 1:
            out = dm_tl_precalc[:, :, i] @ out
        n >>= 1
        i += 1
    return out


def propagate_tb(t_start, t_stop, dt, rho, dm_tl, dm_tl_precalc):
    """timebin_tl.f90:50-77 -- explicit maps while they last, then the stationary fast-forward."""
    n_dm = dm_tl.shape[2]
    n_start = int(_round6(t_start) / dt)
    n_stop = int(_round6(t_stop) / dt)
    n_steps = n_stop - n_start
    steps_dm = min(n_dm - n_start, n_steps)
    v = np.array(rho, dtype=complex)
    while steps_dm > 0:
        v = dm_tl[:, :, n_start] @ v
        n_steps -= 1
        n_start += 1
        steps_dm -= 1
    if n_steps > 0:
        v = fast_propagate(v, dm_tl_precalc, n_steps)
    return v


def apply_left(rho, op, dim):
    return _vec(op @ _mat(rho, dim))


def apply_right(rho, op, dim):
    return _vec(_mat(rho, dim) @ op)


def four_time(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, op_1, op_2, op_3, op_4, tb):
    """timebin_tl.f90:145-214."""
    n_t = len(t1)
    res = np.zeros((n_t, n_t), dtype=complex)
    P = lambda a, b, r, dm: propagate_tb(a, b, dt, r, dm, precalc_tls)
    for i in range(n_t):
        v = P(0.0, t1[i], rho_init, dm_1)
        for j in range(n_t - i):
            t2 = t1[i + j]
            r = apply_right(v, op_1, dim)
            r = P(t1[i], t2, r, dm_1)
            r = apply_right(r, op_2, dim)
            r = P(t2, tb, r, dm_1)
            r = P(0.0, t1[i], r, dm_2)
            r = apply_left(r, op_3, dim)
            r = P(t1[i], t2, r, dm_2)
            r = apply_left(r, op_4, dim)
            res[i, j + i] = _tr(_mat(r, dim))
    return res


def four_time_8op(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, op_et1l, op_et1r, op_et2l, op_et2r,
                  op_lt1l, op_lt1r, op_lt2l, op_lt2r, early_only, late_t1_only, tb):
    """timebin_tl.f90:216-303."""
    n_t = len(t1)
    res = np.zeros((n_t, n_t), dtype=complex)
    P = lambda a, b, r, dm: propagate_tb(a, b, dt, r, dm, precalc_tls)
    for i in range(n_t):
        v = P(0.0, t1[i], rho_init, dm_1)
        for j in range(n_t - i):
            t2 = t1[i + j]
            r = apply_left(apply_right(v, op_et1r, dim), op_et1l, dim)
            r = P(t1[i], t2, r, dm_1)
            r = apply_left(apply_right(r, op_et2r, dim), op_et2l, dim)
            if not early_only:
                r = P(t2, tb, r, dm_1)
                r = P(0.0, t1[i], r, dm_2)
                r = apply_left(apply_right(r, op_lt1r, dim), op_lt1l, dim)
                if not late_t1_only:
                    r = P(t1[i], t2, r, dm_2)
                    r = apply_left(apply_right(r, op_lt2r, dim), op_lt2l, dim)
            res[i, j + i] = _tr(_mat(r, dim))
    return res


def dynamics_t1(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, tb):
    """timebin_tl.f90:305-342."""
    n_t = len(t1)
    res = np.zeros((dim * dim, 2 * n_t - 1), dtype=complex)
    res[:, 0] = rho_init
    for i in range(n_t - 1):
        res[:, i + 1] = propagate_tb(t1[i], t1[i + 1], dt, res[:, i], dm_1, precalc_tls)
    for i in range(n_t - 1):
        res[:, i + n_t] = propagate_tb(t1[i], t1[i + 1], dt, res[:, i + n_t - 1], dm_2, precalc_tls)
    return res


def dynamics_t1_t2(dm_1, dm_2, t1op, t2op, rho_init, t1, precalc_tls, dt, dim, tb, op_1, op_2, op_3):
    """timebin_tl.f90:344-397 (op_2 at t2op overrides op_1 if both times coincide, as in the Fortran)."""
    n_t = len(t1)
    res = np.zeros((dim * dim, 2 * n_t - 1), dtype=complex)
    res[:, 0] = rho_init
    for i in range(n_t - 1):
        r = res[:, i]
        if t1[i] == t1op:
            r = apply_right(res[:, i], op_1, dim)
        if t1[i] == t2op:
            r = apply_right(res[:, i], op_2, dim)
        res[:, i + 1] = propagate_tb(t1[i], t1[i + 1], dt, r, dm_1, precalc_tls)
    for i in range(n_t - 1):
        r = res[:, i + n_t - 1]
        if t1[i] == t1op:
            r = apply_left(res[:, i + n_t - 1], op_3, dim)
        res[:, i + n_t] = propagate_tb(t1[i], t1[i + 1], dt, r, dm_2, precalc_tls)
    return res
