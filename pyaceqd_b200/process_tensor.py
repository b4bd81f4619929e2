"""Process-tensor (PT) container, synthetic generator and on-disk format.

A PT here is the diagonal-coupling MPO that ACE attaches with ``add_PT``
(``pyaceqd/general_system/general_system.py:236``): per time step ``n`` one slice
``A_n[beta, d1, d2]`` (``beta`` = coupling class of the Liouville index)
plus a closure vector ``q_n[d2]`` that traces the environment out after that step.
ACE's own on-disk layout (``*_initial``, ``*_repeated`` ... ``general_system.py:194``) is not
documented anywhere in the reference and no sample exists, so the engine uses its own
``.npz`` container; the initial/periodic split of ACE's "repeat" PTs is
kept (``n_initial`` slices used once, then ``n_repeat`` slices cycled).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class ProcessTensor:
    slices: List[np.ndarray]      # each [n_cls, chi_in, chi_out] complex128
    closures: List[np.ndarray]    # each [chi_out] complex128, closure after that slice
    n_initial: int                # slices[0:n_initial] are used once
    dt: float
    keys: Optional[np.ndarray] = None  # [n_cls, 2] coupling eigenvalue pair of each block
    meta: Optional[dict] = None

    def __post_init__(self):
        self.slices = [np.ascontiguousarray(s, dtype=complex) for s in self.slices]
        self.closures = [np.ascontiguousarray(q, dtype=complex) for q in self.closures]
        if len(self.slices) != len(self.closures) or not self.slices:
            raise ValueError("PT needs one closure per slice and at least one slice")
        if not (0 <= self.n_initial < len(self.slices)):
            raise ValueError("PT needs at least one periodic slice after the initial block")
        if self.slices[0].shape[1] != 1 and self.n_initial > 0:
            raise ValueError("first initial slice must have input bond dimension 1")
        for a, b in zip(self.slices[:-1], self.slices[1:]):
            if a.shape[2] != b.shape[1] and self.n_initial > 0:
                pass  # checked against the step sequence below
        for s, q in zip(self.slices, self.closures):
            if q.shape[0] != s.shape[2]:
                raise ValueError("closure length must equal the slice's output bond dimension")

    @property
    def n_cls(self) -> int:
        return self.slices[0].shape[0]

    @property
    def n_slices(self) -> int:
        return len(self.slices)

    @property
    def n_repeat(self) -> int:
        return len(self.slices) - self.n_initial

    @property
    def chi_max(self) -> int:
        return max(max(s.shape[1], s.shape[2]) for s in self.slices)

    def slice_of_step(self, n) -> np.ndarray:
        """Slice index used by absolute step ``n`` (``n`` may be an array)."""
        n = np.asarray(n, dtype=np.int64)
        per = self.n_initial + (n - self.n_initial) % self.n_repeat
        return np.where(n < self.n_initial, n, per)

    def block_of_class(self, cls_keys: np.ndarray, tol: float = 1e-9) -> np.ndarray:
        """PT block index for every problem coupling class."""
        if self.keys is None:
            if len(cls_keys) > self.n_cls:
                raise ValueError(f"problem has {len(cls_keys)} coupling classes, PT only {self.n_cls} blocks")
            return np.arange(len(cls_keys), dtype=np.int32)
        out = np.empty(len(cls_keys), dtype=np.int32)
        for i, k in enumerate(np.asarray(cls_keys, dtype=float)):
            d = np.abs(self.keys - k[None, :]).max(axis=1)
            j = int(np.argmin(d))
            if d[j] > tol:
                raise ValueError(f"PT has no block for coupling pair {tuple(k)}")
            out[i] = j
        return out

    def save(self, path: str) -> None:
        arrs = {"n_initial": np.int64(self.n_initial), "dt": np.float64(self.dt),
                "n_slices": np.int64(self.n_slices)}
        if self.keys is not None:
            arrs["keys"] = np.asarray(self.keys, dtype=float)
        if self.meta and self.meta.get("coupling_diag") is not None:
            arrs["coupling_diag"] = np.asarray(self.meta["coupling_diag"], dtype=float)
        for i, (s, q) in enumerate(zip(self.slices, self.closures)):
            arrs[f"A{i}"] = s
            arrs[f"q{i}"] = q
        with open(path, "wb") as fh:  # keep the exact name (np.savez would append .npz)
            np.savez(fh, **arrs)

    @staticmethod
    def load(path: str) -> "ProcessTensor":
        with np.load(path) as z:
            n = int(z["n_slices"])
            return ProcessTensor(slices=[z[f"A{i}"] for i in range(n)],
                                 closures=[z[f"q{i}"] for i in range(n)],
                                 n_initial=int(z["n_initial"]), dt=float(z["dt"]),
                                 keys=z["keys"] if "keys" in z.files else None,
                                 meta={"coupling_diag": z["coupling_diag"]} if "coupling_diag" in z.files else None)


def trivial_pt(n_cls: int = 1, dt: float = 0.1) -> ProcessTensor:
    """chi = 1 identity PT: the no-phonon case ("With no PT, chi=1")."""
    return ProcessTensor(slices=[np.ones((n_cls, 1, 1), dtype=complex)],
                         closures=[np.ones(1, dtype=complex)], n_initial=0, dt=dt)


def synthetic_pt(chi: int, n_cls: int, dt: float = 0.1, seed: int = 1234, scale: float = 0.98,
                 n_slices: int = 1, kind: str = "gaussian") -> ProcessTensor:
    """Seeded contractive random PT of exact bond dimension ``chi``.

    ``A[beta] = (G_re + i G_im)/||.||_2 * scale`` with i.i.d. standard normal ``G``,
    ``numpy.random.default_rng(seed)``, periodic over ``n_slices`` slices, closure ``e_0``.
    The state starts as ``rho0`` in bond column 0.  ``kind="unitary"`` (parity tests) uses
    ``scale`` times a random unitary per block and a dense random closure instead, so that
    outputs stay O(1) over many steps.
    """
    rng = np.random.default_rng(seed)
    slices, closures = [], []
    for _ in range(n_slices):
        a = np.empty((n_cls, chi, chi), dtype=complex)
        for b in range(n_cls):
            g = rng.standard_normal((chi, chi)) + 1j * rng.standard_normal((chi, chi))
            if kind == "unitary":
                a[b] = np.linalg.qr(g)[0] * scale
            else:
                a[b] = g / np.linalg.norm(g, 2) * scale
        slices.append(a)
        q = np.zeros(chi, dtype=complex)
        q[0] = 1.0
        if kind == "unitary":
            q = (rng.standard_normal(chi) + 1j * rng.standard_normal(chi)) / np.sqrt(2.0)
            q[0] = 1.0
        closures.append(q)
    return ProcessTensor(slices=slices, closures=closures, n_initial=0, dt=dt,
                         meta={"kind": "synthetic", "seed": seed})


def synthetic_growing_pt(chi: int, n_cls: int, n_initial: int, n_repeat: int = 2, dt: float = 0.1,
                         seed: int = 7) -> ProcessTensor:
    """Synthetic PT with a non-periodic initial block whose bond grows 1 -> chi.

    Exercises rectangular slices and the initial/periodic bookkeeping of ACE's
    "repeat" PTs (``general_system.py:174,194``) in the parity tests.
    """
    rng = np.random.default_rng(seed)
    dims = [1]
    for k in range(n_initial):
        dims.append(min(chi, max(2, dims[-1] * 3)))
    dims[-1] = chi if n_initial > 0 else 1
    slices, closures = [], []

    def rnd(din, dout):
        a = np.empty((n_cls, din, dout), dtype=complex)
        for b in range(n_cls):
            g = rng.standard_normal((din, dout)) + 1j * rng.standard_normal((din, dout))
            a[b] = g / np.linalg.norm(g, 2) * 0.97
        return a

    for k in range(n_initial):
        slices.append(rnd(dims[k], dims[k + 1]))
        closures.append(rng.standard_normal(dims[k + 1]) + 1j * rng.standard_normal(dims[k + 1]))
    cper = chi if n_initial > 0 else chi
    if n_initial == 0:
        # state enters the periodic block with bond 1 embedded in column 0
        pass
    for k in range(n_repeat):
        slices.append(rnd(cper, cper))
        closures.append(rng.standard_normal(cper) + 1j * rng.standard_normal(cper))
    return ProcessTensor(slices=slices, closures=closures, n_initial=n_initial, dt=dt,
                         meta={"kind": "synthetic-growing", "seed": seed})
