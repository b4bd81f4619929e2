"""Numeric problem specification: operator strings -> Liouville-space tensors.

This is the in-process replacement for the param-file the reference writes at
``pyaceqd/general_system/general_system.py:227-290`` and that ACE parses (SURVEY 8a rows
a4/a5).  Conventions:

* density matrix vectorised row-major, ``alpha = nu*N + mu`` for ``|nu><mu|``;
* ``L = -i/hbar [H, .] + sum_k gamma_k (A rho A^+ - 1/2 {A^+A, rho})``;
* ``add_Pulse file F {Op}`` contributes ``f(t)*Op + conj(f(t))*Op^+`` to ``H``
  (reference: ``general_system.py:255,279`` and the comment at ``:247-249``);
* MTO ``_left``: ``rho -> A rho``; ``_right``: ``rho -> rho A`` (A as given);
  ``""``: ``rho -> A rho A^+`` (``general_system.py:29-53``);
* output ``<O> = Tr(O rho)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import constants
from .opparser import parse_operator


def liouville_left(a: np.ndarray) -> np.ndarray:
    """Superoperator of ``rho -> a @ rho`` (row-major vectorisation)."""
    n = a.shape[0]
    return np.kron(a, np.eye(n, dtype=complex))


def liouville_right(a: np.ndarray) -> np.ndarray:
    """Superoperator of ``rho -> rho @ a``."""
    n = a.shape[0]
    return np.kron(np.eye(n, dtype=complex), a.T)


def liouville_sandwich(a: np.ndarray) -> np.ndarray:
    """Superoperator of ``rho -> a @ rho @ a^+``."""
    return np.kron(a, a.conj())


def commutator_generator(h: np.ndarray, hbar: float = constants.hbar) -> np.ndarray:
    """``-i/hbar [h, .]`` as an ``[N^2, N^2]`` matrix."""
    return (-1j / hbar) * (liouville_left(h) - liouville_right(h))


def lindblad_generator(a: np.ndarray, rate: float) -> np.ndarray:
    ada = a.conj().T @ a
    return rate * (liouville_sandwich(a) - 0.5 * liouville_left(ada) - 0.5 * liouville_right(ada))


def output_functional(o: np.ndarray) -> np.ndarray:
    """Row vector ``w`` with ``w @ vec(rho) = Tr(o rho)``."""
    return np.ascontiguousarray(o.T.reshape(-1))


@dataclass
class MTO:
    """One multi-time operator insertion (reference dict at ``general_system.py:29-53``)."""
    superop: np.ndarray  # [NL, NL]
    time: float
    before: bool  # applyBefore == "true": visible already in the output row at ``time``


@dataclass
class Problem:
    """Everything the propagation needs, as arrays.

    ``L0``        [NL, NL]  constant Liouvillian (Hamiltonian + Lindblad terms)
    ``LA, LB``    [n_fields, NL, NL]  ``L(t) = L0 + sum_k f_k(t) LA[k] + conj(f_k(t)) LB[k]``
    ``field_pol`` per field: "x", "y" or "rf" -- which sampled table drives it
    ``rho0``      [NL]      initial vectorised density matrix
    ``out_w``     [n_out, NL]  output functionals
    ``cls``       [NL] int  coupling class of every Liouville index
    ``cls_keys``  [n_cls, 2] (lambda_nu, lambda_mu) of every class
    """
    N: int
    L0: np.ndarray
    LA: np.ndarray
    LB: np.ndarray
    field_pol: List[str]
    rho0: np.ndarray
    out_w: np.ndarray
    cls: np.ndarray
    cls_keys: np.ndarray
    hbar: float = constants.hbar
    H0: Optional[np.ndarray] = None
    meta: Dict[str, object] = field(default_factory=dict)

    @property
    def NL(self) -> int:
        return self.N * self.N

    @property
    def n_out(self) -> int:
        return self.out_w.shape[0]

    @property
    def n_fields(self) -> int:
        return self.LA.shape[0]

    def liouvillian(self, fvals: Sequence[complex]) -> np.ndarray:
        """``L`` for one set of field values (one complex number per field)."""
        L = self.L0.copy()
        for k, f in enumerate(fvals):
            L += f * self.LA[k] + np.conj(f) * self.LB[k]
        return L

    def mto_superop(self, op: np.ndarray, apply_from: str, right_transposed: Optional[bool] = None) -> np.ndarray:
        """``right_transposed`` switches the ``_right`` convention to ``rho -> rho A^T`` (the
        index-order ambiguity of SURVEY App. C.3 / ``dark_model.py:267-268``); default from
        ``constants.mto_right_transposed`` (False: A as given, like ``propagate_tau.f90:91-92``)."""
        if right_transposed is None:
            right_transposed = bool(getattr(constants, "mto_right_transposed", False))
        if apply_from == "_left":
            return liouville_left(op)
        if apply_from == "_right":
            return liouville_right(op.T if right_transposed else op)
        if apply_from == "":
            return liouville_sandwich(op)
        raise ValueError('give "_left" or "_right" or "" for multitime')

    def parse_mtos(self, multitime_op, right_transposed: Optional[bool] = None) -> List[MTO]:
        """Normalise the reference's ``multitime_op`` argument (dict or list of dicts)."""
        if multitime_op is None:
            return []
        if isinstance(multitime_op, dict):
            multitime_op = [multitime_op]
        out = []
        cache = self.meta.setdefault("_mto_cache", {})   # sweeps re-issue the same operators per job
        for m in multitime_op:
            if "operator" not in m or "time" not in m:
                raise ValueError("supply 'operator' and 'time' for multitime")
            key = (m["operator"], m.get("applyFrom", ""), right_transposed,
                   bool(getattr(constants, "mto_right_transposed", False)))
            sup = cache.get(key)
            if sup is None:
                a = parse_operator(m["operator"], self.N)
                sup = cache[key] = self.mto_superop(a, m.get("applyFrom", ""), right_transposed)
            before = str(m.get("applyBefore", "false")).strip().lower() == "true"
            out.append(MTO(sup, float(m["time"]), before))
        return out


def coupling_classes(diag: np.ndarray, tol: float = 1e-12) -> Tuple[np.ndarray, np.ndarray]:
    """Group Liouville indices by the pair of coupling eigenvalues.

    ``diag``: the N diagonal entries of the (diagonal) bath-coupling operator.
    Returns ``cls[NL]`` and ``keys[n_cls, 2]``; classes are numbered in order of first
    appearance so the mapping is deterministic.
    """
    n = len(diag)
    lam = np.real(np.asarray(diag, dtype=complex))
    # snap eigenvalues closer than tol (ACE: dict_zero)
    uniq: List[float] = []
    snapped = np.empty(n)
    for i, v in enumerate(lam):
        for u in uniq:
            if abs(u - v) <= tol:
                snapped[i] = u
                break
        else:
            uniq.append(float(v))
            snapped[i] = float(v)
    keys: List[Tuple[float, float]] = []
    cls = np.empty(n * n, dtype=np.int32)
    for nu in range(n):
        for mu in range(n):
            key = (snapped[nu], snapped[mu])
            if key not in keys:
                keys.append(key)
            cls[nu * n + mu] = keys.index(key)
    return cls, np.asarray(keys, dtype=float).reshape(-1, 2)


def build_problem(*, system_op=None, boson_op=None, initial=None, lindblad_ops=None,
                  interaction_ops=None, output_ops=(), rf_op=None, rho0=None, dim=None,
                  dict_zero: float = 1e-12, polaron_shift: float = 0.0,
                  hbar: float = constants.hbar, raw_pulse_ops=None, coupling_diag=None) -> Problem:
    """Turn the keyword strings of ``system_ace_stream`` into a :class:`Problem`.

    Argument meaning follows ``general_system.py:128-131``:
    ``system_op`` list of Hamiltonian terms; ``lindblad_ops`` list of ``(op, rate)``;
    ``interaction_ops`` list of ``(op, "x"|"y")`` entering as ``-0.5*pi*hbar*(op)`` (``:279``);
    ``rf_op`` entering as ``-0.5*hbar*(rf_op)`` driven by the (real) rf table (``:255``).
    ``polaron_shift``: energy (meV) subtracted as ``-shift * boson_op^2`` -- the
    ``Boson_subtract_polaron_shift`` renormalisation of ``:175`` when a PT is attached.
    ``raw_pulse_ops``: list of ``(op, table)`` taken literally as ``add_Pulse file F {op}`` lines
    (prefactors already inside ``op``; ``table`` in "x", "y", "rf") -- the parameter-file reader.
    ``coupling_diag``: diagonal of the bath coupling operator when it comes with the process tensor
    instead of a ``boson_op`` string.
    """
    # discover the Hilbert dimension from the first operator we can parse
    probe = None
    for cand in ([initial] if initial else []) + list(output_ops or []) + \
            [o[0] for o in (interaction_ops or [])] + [o[0] for o in (raw_pulse_ops or [])] + \
            list(system_op or []) + ([boson_op] if boson_op else []):
        probe = cand
        break
    if dim is None:
        if probe is None:
            if rho0 is None:
                raise ValueError("cannot infer system dimension: no operators given")
            dim = int(np.asarray(rho0).shape[0])
        else:
            dim = parse_operator(probe).shape[0]
    N = int(dim)
    NL = N * N

    H0 = np.zeros((N, N), dtype=complex)
    for s in (system_op or []):
        H0 += parse_operator(s, N)

    coupling_diag_arg, coupling_diag = coupling_diag, np.zeros(N)
    if boson_op is not None:
        bop = parse_operator(boson_op, N)
        off = bop - np.diag(np.diag(bop))
        if np.abs(off).max() > 1e-14:
            raise ValueError("boson_op must be diagonal (diagonal-coupling PT only)")
        coupling_diag = np.real(np.diag(bop))
        if polaron_shift != 0.0:
            H0 = H0 - polaron_shift * np.diag(coupling_diag ** 2)
    elif coupling_diag_arg is not None:
        coupling_diag = np.asarray(coupling_diag_arg, dtype=float).reshape(N)
    cls, keys = coupling_classes(coupling_diag, dict_zero)

    L0 = commutator_generator(H0, hbar)
    for op, rate in (lindblad_ops or []):
        L0 += lindblad_generator(parse_operator(op, N), float(rate))

    LA, LB, pol = [], [], []
    if rf_op is not None:
        a = parse_operator("-0.5*hbar*({})".format(rf_op), N)
        LA.append(commutator_generator(a, hbar))
        LB.append(commutator_generator(a.conj().T, hbar))
        pol.append("rf")
    for op, p in (interaction_ops or []):
        a = parse_operator("-0.5*pi*hbar*({})".format(op), N)
        LA.append(commutator_generator(a, hbar))
        LB.append(commutator_generator(a.conj().T, hbar))
        pol.append("y" if p == "y" else "x")
    for op, table in (raw_pulse_ops or []):
        a = parse_operator(op, N)
        LA.append(commutator_generator(a, hbar))
        LB.append(commutator_generator(a.conj().T, hbar))
        pol.append(table)
    LA = np.asarray(LA, dtype=complex).reshape(-1, NL, NL)
    LB = np.asarray(LB, dtype=complex).reshape(-1, NL, NL)

    if rho0 is not None:
        r0 = np.asarray(rho0, dtype=complex).reshape(N, N)
    elif initial is not None:
        r0 = parse_operator(initial, N)
    else:
        r0 = np.zeros((N, N), dtype=complex)
        r0[0, 0] = 1.0
    out_w = np.asarray([output_functional(parse_operator(o, N)) for o in output_ops],
                       dtype=complex).reshape(-1, NL)
    return Problem(N=N, L0=np.ascontiguousarray(L0), LA=LA, LB=LB, field_pol=pol,
                   rho0=np.ascontiguousarray(r0.reshape(-1)), out_w=out_w, cls=cls,
                   cls_keys=keys, hbar=hbar, H0=H0,
                   meta={"coupling_diag": coupling_diag})
