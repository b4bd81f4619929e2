"""CPU ORACLE (test infrastructure, NOT a product path) -- NumPy restatement of the hot path.

PARITY UNPINNED: the arithmetic of this path lives in the external ACE code
(github.com/mcygorek/ACE, no version pinned by the reference: ``setup.py:10-17``,
``README.md:4-8``); ACE is absent from /root/reference and from this container, and the
reference holds no numeric golden vector for any propagation result.  This file
restates the published algorithm (Cygorek et al., Nat. Phys. 18, 662 (2022))
as it is driven by the reference's call site ``pyaceqd/general_system/general_system.py:227-290``
and is pinned by physics known-answer tests (tests/test_oracle_kat.py) and, for the coherent part of
the path, at plot resolution (0.005) by the one ACE-made artefact the reference keeps, the plot
``pyaceqd/tests/sixls_compare.png`` (tests/test_reference_plot.py; the phonon part stays unpinned).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (pyaceqd_b200) never does.

Objects are duck-typed (``problem``: L0, LA, LB, field_pol, rho0, out_w, cls, cls_keys;
``pt``: slices, closures, n_initial, slice_of_step, block_of_class; ``job``: t_start, t_end,
dt, tables, mtos) so the oracle does not import the product package.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

try:   # small matrix-vector products: a threaded BLAS spends its time spinning (measured 40 ms -> 0.2 ms per slice)
    from threadpoolctl import threadpool_limits
except ImportError:   # pragma: no cover
    from contextlib import contextmanager

    @contextmanager
    def threadpool_limits(limits=None):
        yield

T_EVAL_CHOICES = ("half_mid", "step_mid", "start")


def n_steps_of(job) -> int:
    """ACE time grid: N = round((te-ta)/dt) steps, N+1 output rows (param keys written at
    general_system.py:229-231)."""
    return int(round((job.t_end - job.t_start) / job.dt))


def sample_field(table, t: float, n_valid: int = 0) -> complex:
    """Value of a tabulated drive at time ``t``: linear interpolation between the samples of
    the pulse file (general_system.py:55-71 writes them on ``t0 + j*dt``), end values held
    outside the table.  ``n_valid`` > 0: only the first ``n_valid`` samples belong to this run's
    own pulse file (the reference writes it on np.arange(t_start, t_end, dt), :213)."""
    x = (t - table.t0) / table.dt
    v = table.values
    n = len(v) if n_valid <= 0 else min(len(v), n_valid)
    if n == 0:
        return 0.0 + 0.0j
    if x <= 0:
        return complex(v[0])
    if x >= n - 1:
        return complex(v[n - 1])
    j = int(np.floor(x))
    w = x - j
    return complex((1.0 - w) * v[j] + w * v[j + 1])


def liouvillian_at(problem, job, t: float) -> np.ndarray:
    """L(t) = L0 + sum_k f_k(t) LA_k + conj(f_k(t)) LB_k   (add_Pulse semantics,
    general_system.py:255,279)."""
    L = np.array(problem.L0, dtype=complex, copy=True)
    for k, pol in enumerate(problem.field_pol):
        tab = job.tables.get(pol)
        if tab is None:
            continue
        f = sample_field(tab, t, getattr(job, "table_len", 0))
        L += f * problem.LA[k] + np.conj(f) * problem.LB[k]
    return L


def half_step_times(t_n: float, dt: float, t_eval: str):
    if t_eval == "half_mid":
        return t_n + 0.25 * dt, t_n + 0.75 * dt
    if t_eval == "step_mid":
        return t_n + 0.5 * dt, t_n + 0.5 * dt
    if t_eval == "start":
        return t_n, t_n + 0.5 * dt
    raise ValueError(t_eval)


def apply_slice(state: np.ndarray, A: np.ndarray, blk_of_alpha: np.ndarray) -> np.ndarray:
    """state'[alpha, d2] = sum_d1 A[beta(alpha), d1, d2] state[alpha, d1]   (SURVEY 8a row a6;
    diagonal coupling).  A state narrower than the slice input is zero-extended (a fresh
    state entering a periodic block sits in bond column 0)."""
    NL, chi = state.shape
    din, dout = A.shape[1], A.shape[2]
    if chi < din:
        state = np.concatenate([state, np.zeros((NL, din - chi), dtype=complex)], axis=1)
    elif chi > din:
        raise ValueError("state bond dimension exceeds slice input dimension")
    out = np.empty((NL, dout), dtype=complex)
    for a in range(NL):
        out[a] = state[a] @ A[blk_of_alpha[a]]
    return out


def propagate(problem, pt, job, t_eval: str = "half_mid", return_states: bool = False):
    """One trajectory, literally SURVEY App. D.2 (symmetric Trotter, use_symmetric_Trotter
    true at general_system.py:234):

        state <- M(t_n, dt/2) state ; PT slice n ; state <- M(t_n+dt/2, dt/2) state
        MTOs scheduled at t_{n+1} (file order; applyBefore ones before the output row)
        rho(t_{n+1}) = state . q_{n+1} ; out_j = w_j . rho

    Returns ``out[n_out, N+1]`` complex (row k at time t_start + k dt).
    """
    with threadpool_limits(limits=1):
        return _propagate(problem, pt, job, t_eval, return_states)


def _propagate(problem, pt, job, t_eval, return_states):
    N = n_steps_of(job)
    NL = problem.L0.shape[0]
    blk_of_cls = pt.block_of_class(problem.cls_keys)
    blk_of_alpha = blk_of_cls[np.asarray(problem.cls)]
    out = np.zeros((problem.out_w.shape[0], N + 1), dtype=complex)

    mto_at = {}
    for m in job.mtos:
        k = int(round((m.time - job.t_start) / job.dt))
        if k < 0 or k > N:
            raise ValueError("multitime operator outside the time window")
        mto_at.setdefault(k, []).append(m)

    state = np.zeros((NL, 1), dtype=complex)
    state[:, 0] = problem.rho0
    closure = np.ones(1, dtype=complex)
    states = []

    def row(k):
        nonlocal state
        for m in mto_at.get(k, []):
            if m.before:
                state = m.superop @ state
        rho = state[:, :len(closure)] @ closure
        out[:, k] = problem.out_w @ rho
        for m in mto_at.get(k, []):
            if not m.before:
                state = m.superop @ state

    row(0)
    for n in range(N):
        t_n = job.t_start + n * job.dt
        ta, tb = half_step_times(t_n, job.dt, t_eval)
        M1 = expm(liouvillian_at(problem, job, ta) * (0.5 * job.dt))
        M2 = expm(liouvillian_at(problem, job, tb) * (0.5 * job.dt))
        state = M1 @ state
        s = int(pt.slice_of_step(n))
        state = apply_slice(state, pt.slices[s], blk_of_alpha)
        state = M2 @ state
        closure = pt.closures[s]
        if return_states:
            states.append(state.copy())
        row(n + 1)
    if return_states:
        return out, states
    return out


def dynamical_map(problem, pt, job, t_eval: str = "half_mid") -> np.ndarray:
    """E[t_i] (``DynamicalMap.E`` consumed at general_system.py:328-335): column j of E_i is
    rho(t_i) for the j-th unit vector as initial state.  Shape (N+1, NL, NL)."""
    NL = problem.L0.shape[0]
    N = n_steps_of(job)
    E = np.zeros((N + 1, NL, NL), dtype=complex)

    class _P:
        pass

    for j in range(NL):
        p = _P()
        p.__dict__.update({k: getattr(problem, k) for k in
                           ("L0", "LA", "LB", "field_pol", "cls", "cls_keys")})
        p.rho0 = np.zeros(NL, dtype=complex)
        p.rho0[j] = 1.0
        p.out_w = np.eye(NL, dtype=complex)
        E[:, :, j] = propagate(p, pt, job, t_eval).T
    return E
