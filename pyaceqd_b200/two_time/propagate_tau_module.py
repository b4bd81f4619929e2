"""GPU stand-in for the reference's f2py module ``propagate_tau_module``
(``pyaceqd/two_time/propagate_tau.f90``, built with ``f2py -c --f90flags="-fopenmp" ... -lopenblas``,
``:1``).  Same routine names, argument order and array layouts as the f2py wrappers the reference
calls (``two_time/correlations.py:534,583,782,831``, ``two_time/purity.py:602,709,741,770``): maps
as ``[NL, NL, n]``, operators ``[dim, dim]``, results ``[n_t, n_tau + 1]``.

The sequential prefix (propagating rho to every ``t`` of ``time_sparse``) is a short chain and
stays on the host; the O(n_t * n_tau) matrix-vector chains run as one launch of the chain kernel
(csrc/tlmap.cu).  Column-major reshapes of the Fortran side are kept (see pyaceqd_b200/tlmap.py).
"""
from __future__ import annotations

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.tlmap import Programs, left_superop, maps_first, right_superop, rle, trace_functional


def _mat(v, dim):
    return np.asarray(v).reshape(dim, dim, order="F")


def propagate_tau(dm_tl, rho_init, n_tau, dim, j_start):
    """``rho_out[:, k] = dm_tl[:, :, j_start + k - 1] rho_out[:, k - 1]`` (propagate_tau.f90:3-19)."""
    NL = dim * dim
    pool = maps_first(dm_tl)
    pr = Programs(NL)
    off = pr.add(pool)
    pr.chain(rho_init, [(off + j_start, n_tau, 1, 1)])
    out, _ = pr.run(w=np.eye(NL, dtype=complex))
    res = np.empty((NL, n_tau + 1), dtype=complex, order="F")
    res[:, 0] = rho_init
    res[:, 1:] = out[0, :n_tau].T
    return res


def _prefix(rho_init, time, time_sparse, map_of_step):
    """States and step counters at every sparse time (the serial part of every calc_* routine)."""
    v = np.array(rho_init, dtype=complex)
    j, states, js = 0, [], []
    for ts in time_sparse:
        while time[j] < ts:
            v = map_of_step(j) @ v
            j += 1
        states.append(v.copy())
        js.append(j)
    return np.asarray(states), np.asarray(js, dtype=np.int64)


def calc_onetime_parallel(dm_tl, rho_init, n_tau, dim, opa, opb, opc, time, time_sparse):
    """``result[i, 0] = Tr(A B C rho(t_i))``; ``result[i, k] = Tr(B rho_k)`` with ``rho_0 = C rho(t_i) A``
    pushed through ``dm_tl[j_i], dm_tl[j_i + 1], ...`` (propagate_tau.f90:110-187)."""
    NL = dim * dim
    pool = maps_first(dm_tl)
    states, js = _prefix(rho_init, time, time_sparse, lambda j: pool[j])
    abc = trace_functional(opa @ opb @ opc)
    start = right_superop(np.asarray(opa, dtype=complex)) @ left_superop(np.asarray(opc, dtype=complex))
    pr = Programs(NL)
    off = pr.add(pool)
    for v, j in zip(states, js):
        pr.chain(start @ v, [(off + int(j), n_tau, 1, 1)])
    out, _ = pr.run(w=trace_functional(np.asarray(opb, dtype=complex))[None])
    res = np.empty((len(time_sparse), n_tau + 1), dtype=complex)
    res[:, 0] = states @ abc
    res[:, 1:] = out[:, :n_tau, 0]
    return res


calc_onetime = calc_onetime_parallel


def _periodic_indices(j0, K, n_tb, n_map, stationary):
    """Matrix indices of K steps starting at period position ``j0`` (1-based like the Fortran): inside
    a period the first n_map steps use block maps 0..n_map-1, the rest the stationary map."""
    pos = (j0 - 1 + np.arange(K)) % n_tb + 1 if j0 <= n_tb else None
    if pos is None:      # started beyond the first period: the Fortran counter only wraps at n_tb + 1
        pos = j0 + np.arange(K)
    return np.where(pos <= n_map, pos - 1, stationary)


def calc_onetime_parallel_block(dm_block, dm_s, rho_init, n_tb, nx_tau, dim, opa, opb, opc, time, time_sparse):
    """Periodic maps: ``dm_block`` for the first ``n_map`` steps of each period of ``n_tb`` steps, then
    the stationary ``dm_s`` (propagate_tau.f90:189-295)."""
    NL = dim * dim
    block = maps_first(dm_block)
    n_map = block.shape[0]
    dm_s = np.asarray(dm_s, dtype=complex)
    states, js = _prefix(rho_init, time, time_sparse, lambda j: block[j] if j < n_map else dm_s)
    K = nx_tau * n_tb
    abc = trace_functional(opa @ opb @ opc)
    start = right_superop(np.asarray(opa, dtype=complex)) @ left_superop(np.asarray(opc, dtype=complex))
    pr = Programs(NL)
    off = pr.add(block)
    i_s = pr.add(dm_s)
    for v, j in zip(states, js):
        idx = _periodic_indices(int(j) + 1, K, n_tb, n_map, i_s - off) + off
        pr.chain(start @ v, rle(idx, True))
    out, _ = pr.run(w=trace_functional(np.asarray(opb, dtype=complex))[None])
    res = np.empty((len(time_sparse), K + 1), dtype=complex)
    res[:, 0] = states @ abc
    res[:, 1:] = out[:, :K, 0]
    return res


def calc_twotime_phonon_block(dm_taucs2, dm_sep1, dm_sep2, dm_s, rho_init, n_tb, nx_tau, dim, opa, opb, opc,
                              time, time_sparse):
    """Phonon-aware variant (propagate_tau.f90:374-536): until the end of the period in which the
    operators act, chain ``i`` uses its own maps ``dm_taucs2[:, :, i, :]`` (first ``n_tauc`` chains) or
    ``dm_sep2`` (the others); afterwards ``dm_sep1``.  The branch starts from ``rho(t_i)`` unchanged
    (:455-459) and the trace is taken with ``transpose(opB)`` (:484), exactly as in the Fortran."""
    NL = dim * dim
    sep1, sep2 = maps_first(dm_sep1), maps_first(dm_sep2)
    n_map = sep1.shape[0]
    dm_s = np.asarray(dm_s, dtype=complex)
    taucs = np.asarray(dm_taucs2, dtype=complex)            # [NL, NL, n_tauc, n_map]
    n_tauc = taucs.shape[2]
    states, js = _prefix(rho_init, time, time_sparse, lambda j: sep1[j] if j < n_map else dm_s)
    K = nx_tau * n_tb
    pr = Programs(NL)
    o1, o2, o_s = pr.add(sep1), pr.add(sep2), pr.add(dm_s)
    o_t = pr.add(np.ascontiguousarray(np.transpose(taucs, (2, 3, 0, 1)).reshape(n_tauc * n_map, NL, NL))) \
        if n_tauc else 0
    steps = np.arange(K)
    align = 1 if constants.phonon_block_aligned else 0      # see constants.py: the Fortran resets one step early
    for i, (v, j) in enumerate(zip(states, js)):
        j1 = int(j) + 1                                     # Fortran's 1-based j_array(i)
        first = n_tb - j1 + align                           # steps before `j + j_start == n_tb + 1` fires
        if first >= 1:
            pos = np.where(steps < first, steps + 1, (steps - first) % n_tb + 1)
            own = steps < first
        else:                                               # the reset never fires (:488-492)
            pos, own = steps + 1, np.ones(K, dtype=bool)
        base2 = (o_t + i * n_map) if i < n_tauc else o2
        idx = np.where(pos <= n_map, pos - 1 + np.where(own, base2, o1), o_s)
        pr.chain(v, rle(idx, True))
    out, _ = pr.run(w=trace_functional(np.asarray(opb, dtype=complex).T)[None])
    res = np.empty((len(time_sparse), K + 1), dtype=complex)
    res[:, 0] = states @ trace_functional(opa @ opb @ opc)
    res[:, 1:] = out[:, :K, 0]
    return res
