"""Host side of the time-local-map chain kernel (``aceqd_tlmap_run``, csrc/tlmap.cu).

The reference's Fortran helpers (``two_time/propagate_tau.f90``, ``timebin/timebin_tl.f90``) are
all "push a Liouville vector through a schedule of NL x NL matrices, insert operators, take
traces".  Here such a schedule is a *program* over one matrix pool; this module builds programs
(run-length encoding of matrix-index sequences into segments) and holds the column-major
conventions of the Fortran side:

    Fortran ``reshape(v, [dim, dim])`` is column-major, so for ``M = mat(v)``
      apply_left  : vec(op M)  = kron(I, op)   v          (timebin_tl.f90:101-110)
      apply_right : vec(M op)  = kron(op^T, I) v          (timebin_tl.f90:112-121)
      Tr(B M)     = B.reshape(-1) . v                      (propagate_tau.f90:101-103)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

import pyaceqd_b200.engine as _engine
from pyaceqd_b200.engine import TLSEG_DT


def left_superop(op: np.ndarray) -> np.ndarray:
    return np.kron(np.eye(op.shape[0]), op)


def right_superop(op: np.ndarray) -> np.ndarray:
    return np.kron(op.T, np.eye(op.shape[0]))


def trace_functional(op: np.ndarray) -> np.ndarray:
    """``w`` with ``w . v = Tr(op mat(v))``."""
    return np.asarray(op, dtype=complex).reshape(-1)


def maps_first(a: np.ndarray) -> np.ndarray:
    """f2py layout ``[NL, NL, n]`` -> pool layout ``[n, NL, NL]``."""
    return np.ascontiguousarray(np.moveaxis(np.asarray(a, dtype=complex), 2, 0))


def rle(idx: np.ndarray, emit: bool) -> List[Tuple[int, int, int, int]]:
    """Run-length encode one chain's matrix-index sequence into ``(start, count, emit, stride)``
    segments: maximal runs of consecutive (+1) or repeated (+0) indices."""
    idx = np.asarray(idx, dtype=np.int64)
    n = len(idx)
    if n == 0:
        return []
    d = np.diff(idx)
    segs, a = [], 0
    while a < n:
        if a + 1 < n and d[a] in (0, 1):
            step = int(d[a])
            b = a + 1
            while b < n - 1 and d[b] == step:
                b += 1
            segs.append((int(idx[a]), b - a + 1, int(emit), step))
            a = b + 1
        else:
            segs.append((int(idx[a]), 1, int(emit), 1))
            a += 1
    return segs


class Programs:
    """Matrix pool + per-chain segment lists + start vectors for one ``aceqd_tlmap_run`` call."""

    def __init__(self, NL: int):
        self.NL = NL
        self._mats: List[np.ndarray] = []
        self._n = 0
        self.chains: List[List[Tuple[int, int, int, int]]] = []
        self.v0: List[np.ndarray] = []

    def add(self, mats: np.ndarray) -> int:
        """Append ``[n, NL, NL]`` (or one ``[NL, NL]``) to the pool; returns the first index."""
        m = np.asarray(mats, dtype=complex)
        if m.ndim == 2:
            m = m[None]
        off = self._n
        self._mats.append(np.ascontiguousarray(m))
        self._n += m.shape[0]
        return off

    def chain(self, v0: np.ndarray, segs: Sequence[Tuple[int, int, int, int]]):
        self.v0.append(np.asarray(v0, dtype=complex))
        self.chains.append([s for s in segs if s[1] > 0])

    def run(self, w=None, want_final=False, engine=None):
        eng = engine or _engine.default_engine()
        seg_off = np.zeros(len(self.chains) + 1, dtype=np.int64)
        seg_off[1:] = np.cumsum([len(c) for c in self.chains])
        segs = np.zeros(int(seg_off[-1]), dtype=TLSEG_DT)
        flat = [s for c in self.chains for s in c]
        if flat:
            arr = np.asarray(flat, dtype=np.int32)
            segs["start"], segs["count"], segs["emit"], segs["stride"] = arr.T
        n_emit = max([sum(s[1] for s in c if s[2]) for c in self.chains] + [0])
        pool = np.concatenate(self._mats, axis=0)
        return eng.tlmap_run(pool, np.asarray(self.v0).reshape(len(self.chains), self.NL), seg_off, segs,
                             w=w, n_emit_max=n_emit, want_final=want_final)
