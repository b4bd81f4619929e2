"""GPU stand-in for the reference's f2py module ``timebin_tl`` (``pyaceqd/timebin/timebin_tl.f90``).

Same routine names, argument order and layouts as the f2py wrappers the reference calls
(``timebin/twophoton_new.py:661,697,712,757,835-841,915``): ``four_time``, ``four_time_8op``,
``dynamics_t1``, ``dynamics_t1_t2`` and the ``utils`` namespace (``fast_propagate``,
``propagate_tb``, ``apply_left``, ``apply_right``, ...).

The O(n_t^2) double loops of ``four_time`` / ``four_time_8op`` (``:145-303``) become one launch of the
chain kernel: every pair ``t1 <= t2`` is a program over a pool holding both bins' time-local maps,
the binary powers of the stationary map and the operator superoperators.  The short sequential
utilities stay on the host.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from pyaceqd_b200.tlmap import Programs, left_superop, maps_first, right_superop, trace_functional


def _round6(x):
    return np.rint(x * 1_000_000) / 1_000_000


def _mat(v, dim):
    return np.asarray(v).reshape(dim, dim, order="F")


def _vec(m):
    return np.asarray(m).reshape(-1, order="F")


# ------------------------------------------------------------------ utils (host, sequential)
def fast_propagate(rho, dm_tl_precalc, n_steps):
    """``E^n rho`` from precomputed ``E^(2^i)`` (timebin_tl.f90:23-47)."""
    out = np.array(rho, dtype=complex)
    i, n = 0, int(n_steps)
    while n > 0:
        if n & 1:
            out = dm_tl_precalc[:, :, i] @ out
        n >>= 1
        i += 1
    return out


def _tb_schedule(t_start, t_stop, dt, n_dm):
    """(first explicit map, number of explicit maps, stationary steps) of one ``propagate_tb``."""
    n_start = int(_round6(t_start) / dt)
    n_steps = int(_round6(t_stop) / dt) - n_start
    explicit = max(0, min(n_dm - n_start, n_steps))
    return n_start, explicit, max(0, n_steps - explicit)


def propagate_tb(t_start, t_stop, dt, rho, dm_tl, dm_tl_precalc):
    """Explicit time-local maps while they last, then the stationary fast-forward (timebin_tl.f90:50-77)."""
    n0, explicit, rest = _tb_schedule(t_start, t_stop, dt, dm_tl.shape[2])
    v = np.array(rho, dtype=complex)
    for k in range(explicit):
        v = dm_tl[:, :, n0 + k] @ v
    return fast_propagate(v, dm_tl_precalc, rest) if rest > 0 else v


def apply_left(rho, op, dim):
    return _vec(np.asarray(op) @ _mat(rho, dim))


def apply_right(rho, op, dim):
    return _vec(_mat(rho, dim) @ np.asarray(op))


def apply_matrix_from_left(rho, op, dim):
    return np.asarray(op) @ np.asarray(rho)


def apply_matrix_from_right(rho, op, dim):
    return np.asarray(rho) @ np.asarray(op)


def apply_operator(rho, op, dim):
    return np.asarray(op) @ np.asarray(rho)


def test_reshape(rho, dim):
    return np.asarray(rho).copy()


utils = SimpleNamespace(fast_propagate=fast_propagate, propagate_tb=propagate_tb, apply_left=apply_left,
                        apply_right=apply_right, apply_matrix_from_left=apply_matrix_from_left,
                        apply_matrix_from_right=apply_matrix_from_right, apply_operator=apply_operator,
                        test_reshape=test_reshape)


# ------------------------------------------------------------------ batched double loops (device)
class _Pool:
    """Both bins' maps + stationary powers + operator superoperators in one program pool."""

    def __init__(self, dm_1, dm_2, precalc_tls, dim):
        self.pr = Programs(dim * dim)
        self.n_dm = np.asarray(dm_1).shape[2]
        self.off = {1: self.pr.add(maps_first(dm_1)), 2: self.pr.add(maps_first(dm_2))}
        self.pre = self.pr.add(maps_first(precalc_tls))

    def op(self, matrix):
        return [(self.pr.add(matrix), 1, 0, 1)]

    def tb(self, which, t_start, t_stop, dt):
        n0, explicit, rest = _tb_schedule(t_start, t_stop, dt, self.n_dm)
        segs = [(self.off[which] + n0, explicit, 0, 1)] if explicit else []
        i = 0
        while rest > 0:
            if rest & 1:
                segs.append((self.pre + i, 1, 0, 1))
            rest >>= 1
            i += 1
        return segs


def _emit_last(segs):
    s = list(segs)
    st, cnt, _, stride = s[-1]
    if cnt > 1:           # only the very last step emits
        s[-1] = (st, cnt - 1, 0, stride)
        s.append((st + (cnt - 1) * stride, 1, 1, stride))
    else:
        s[-1] = (st, 1, 1, stride)
    return s


def _pairs(pool, rho_init, t1, dt, dim, tb, stages, dm_1, precalc_tls):
    """``result[i, i + j]`` = trace after running, for every pair ``t1[i] <= t2 = t1[i + j]``, the
    stage list ``stages(t1_i, t2)`` from ``rho(t1_i)``."""
    n_t = len(t1)
    res = np.zeros((n_t, n_t), dtype=complex)
    index = []
    for i in range(n_t):
        v = propagate_tb(0.0, t1[i], dt, rho_init, dm_1, precalc_tls)     # serial prefix, O(n_t) chains
        for j in range(n_t - i):
            segs = stages(t1[i], t1[i + j])
            pool.pr.chain(v, _emit_last(segs))
            index.append((i, i + j))
    out, _ = pool.pr.run(w=trace_functional(np.eye(dim))[None])
    for (i, k), val in zip(index, out[:, 0, 0]):
        res[i, k] = val
    return res


def four_time(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, op_1, op_2, op_3, op_4, tb):
    """timebin_tl.f90:145-214: ``op_1`` (right) at t1, ``op_2`` (right) at t2 in the early bin, to the end
    of the bin, then ``op_3`` (left) at t1 and ``op_4`` (left) at t2 in the late bin; trace."""
    pool = _Pool(dm_1, dm_2, precalc_tls, dim)
    o1, o2 = pool.op(right_superop(np.asarray(op_1, complex))), pool.op(right_superop(np.asarray(op_2, complex)))
    o3, o4 = pool.op(left_superop(np.asarray(op_3, complex))), pool.op(left_superop(np.asarray(op_4, complex)))

    def stages(ta, tb2):
        return (o1 + pool.tb(1, ta, tb2, dt) + o2 + pool.tb(1, tb2, tb, dt) + pool.tb(2, 0.0, ta, dt) + o3 +
                pool.tb(2, ta, tb2, dt) + o4)
    return _pairs(pool, rho_init, t1, dt, dim, tb, stages, np.asarray(dm_1), np.asarray(precalc_tls))


def four_time_8op(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, op_et1l, op_et1r, op_et2l, op_et2r,
                  op_lt1l, op_lt1r, op_lt2l, op_lt2r, early_only, late_t1_only, tb):
    """timebin_tl.f90:216-303: a (right, then left) operator pair at each of the four times; ``early_only``
    / ``late_t1_only`` stop after the second / third pair."""
    pool = _Pool(dm_1, dm_2, precalc_tls, dim)
    c = lambda o: np.asarray(o, dtype=complex)
    pair = lambda left, right: pool.op(left_superop(c(left)) @ right_superop(c(right)))
    e1, e2 = pair(op_et1l, op_et1r), pair(op_et2l, op_et2r)
    l1, l2 = pair(op_lt1l, op_lt1r), pair(op_lt2l, op_lt2r)

    def stages(ta, tb2):
        s = e1 + pool.tb(1, ta, tb2, dt) + e2
        if early_only:
            return s
        s = s + pool.tb(1, tb2, tb, dt) + pool.tb(2, 0.0, ta, dt) + l1
        if late_t1_only:
            return s
        return s + pool.tb(2, ta, tb2, dt) + l2
    return _pairs(pool, rho_init, t1, dt, dim, tb, stages, np.asarray(dm_1), np.asarray(precalc_tls))


def dynamics_t1(dm_1, dm_2, rho_init, t1, precalc_tls, dt, dim, tb):
    """Density vectors on the ``t1`` grid through both bins (timebin_tl.f90:305-342); serial, host."""
    n_t = len(t1)
    res = np.zeros((dim * dim, 2 * n_t - 1), dtype=complex, order="F")
    res[:, 0] = rho_init
    for b, dm in enumerate((dm_1, dm_2)):
        for i in range(n_t - 1):
            k = i + b * (n_t - 1)
            res[:, k + 1] = propagate_tb(t1[i], t1[i + 1], dt, res[:, k], dm, precalc_tls)
    return res


def dynamics_t1_t2(dm_1, dm_2, t1op, t2op, rho_init, t1, precalc_tls, dt, dim, tb, op_1, op_2, op_3):
    """As :func:`dynamics_t1` with ``op_1`` / ``op_2`` (right) at ``t1op`` / ``t2op`` in the early bin and
    ``op_3`` (left) at ``t1op`` in the late bin (timebin_tl.f90:344-397)."""
    n_t = len(t1)
    res = np.zeros((dim * dim, 2 * n_t - 1), dtype=complex, order="F")
    res[:, 0] = rho_init
    for i in range(n_t - 1):
        r = res[:, i]
        if t1[i] == t1op:
            r = apply_right(res[:, i], op_1, dim)
        if t1[i] == t2op:
            r = apply_right(res[:, i], op_2, dim)
        res[:, i + 1] = propagate_tb(t1[i], t1[i + 1], dt, r, dm_1, precalc_tls)
    for i in range(n_t - 1):
        k = i + n_t - 1
        r = res[:, k]
        if t1[i] == t1op:
            r = apply_left(res[:, k], op_3, dim)
        res[:, k + 1] = propagate_tb(t1[i], t1[i + 1], dt, r, dm_2, precalc_tls)
    return res
