"""Process-tensor construction for a Gaussian bosonic bath with diagonal coupling.

Replaces the PT-generation run the reference delegates to ACE
(``pyaceqd/general_system/general_system.py:159-192``: ``Boson_SysOp``, ``Boson_J_type QDPhonon``,
``Boson_J_a_e/a_h``, ``temperature``, ``threshold``, ``Boson_subtract_polaron_shift``,
``use_Gaussian_infinite`` ...).  ACE's source is not available here, so this follows the published
algorithms:

* spectral density of the deformation-potential coupling of a GaAs quantum dot to LA phonons,
  ``J(w) = w^3/(4 pi^2 rho hbar c_s^5) (D_e e^{-w^2 a_e^2/4c_s^2} - D_h e^{-w^2 a_h^2/4c_s^2})^2``;
* discretised influence functional ``F = prod_n prod_{k>=0} I_k(c_n, c_{n-k})`` with
  ``I_k(c, c') = exp(-(l+ - l-)(eta_k l'+ - conj(eta_k) l'-))`` and the QUAPI coefficients
  ``eta_k`` (Makri & Makarov 1995; Strathearn et al. 2018);
* the uniform ("infinite") process tensor of that functional by iTEBD contraction of the
  time-translation-invariant network (Link, Tu, Strunz, PRL 132, 200403 (2024)) -- the form ACE
  produces with ``use_Gaussian_infinite`` (``general_system.py:165-167``); the finite-memory
  "repeat" form of ``:169-174`` describes the same functional, so both reference modes map to
  this one construction (parity with ACE is at truncation level by nature).

The result is a :class:`~pyaceqd_b200.process_tensor.ProcessTensor` with a single periodic slice,
gauge-fixed so that the initial bond state is ``e_0``.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import constants
from .problem import coupling_classes
from .process_tensor import ProcessTensor

# GaAs material parameters
RHO = 5370.0           # kg/m^3
C_S = 5110.0           # m/s
D_E = 7.0              # eV
D_H = -3.5             # eV
_HBAR_SI = 1.054571817e-34
_EV = 1.602176634e-19


def qd_phonon_spectral_density(w, a_e: float = 5.0, a_h: Optional[float] = None) -> np.ndarray:
    """``J(w)`` in 1/ps for angular frequency ``w`` in 1/ps; ``a_e``, ``a_h`` in nm
    (``a_h = a_e/1.15`` by default, reference ``rabi_rotations.py:17``)."""
    if a_h is None:
        a_h = a_e / 1.15
    w_si = np.asarray(w, dtype=float) * 1e12
    ae, ah = a_e * 1e-9, a_h * 1e-9
    form = D_E * _EV * np.exp(-(w_si * ae) ** 2 / (4 * C_S ** 2)) - D_H * _EV * np.exp(-(w_si * ah) ** 2 / (4 * C_S ** 2))
    j_si = w_si ** 3 / (4 * np.pi ** 2 * RHO * _HBAR_SI * C_S ** 5) * form ** 2    # 1/s
    return j_si * 1e-12


def eta_coefficients(J: np.ndarray, w: np.ndarray, dt: float, K: int, temperature: float) -> np.ndarray:
    """QUAPI influence coefficients ``eta_0 .. eta_K`` (dimensionless) for a bath correlation
    function ``C(t) = int dw J(w) [coth(hbar w / 2 k_B T) cos wt - i sin wt]``:

        eta_0 = int dw J/w^2 [coth (1 - cos w dt) + i (sin w dt - w dt)]
        eta_k = int dw J/w^2  2 (1 - cos w dt) [coth cos(k w dt) - i sin(k w dt)],  k >= 1
    """
    w = np.asarray(w, dtype=float)
    J = np.asarray(J, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if temperature > 0:
            coth = 1.0 / np.tanh(constants.hbar * w / (2 * constants.kB * temperature))
        else:
            coth = np.ones_like(w)
        jw2 = np.where(w > 0, J / w ** 2, 0.0)
        jc = np.where(w > 0, jw2 * coth, 0.0)
    jc = np.nan_to_num(jc, nan=0.0, posinf=0.0)
    eta = np.empty(K + 1, dtype=complex)
    wd = w * dt
    eta[0] = np.trapezoid(jc * (1 - np.cos(wd)), w) + 1j * np.trapezoid(jw2 * (np.sin(wd) - wd), w)
    one_m_cos = 2 * (1 - np.cos(wd))
    for k in range(1, K + 1):
        eta[k] = np.trapezoid(jc * one_m_cos * np.cos(k * wd), w) - 1j * np.trapezoid(jw2 * one_m_cos * np.sin(k * wd), w)
    return eta


def polaron_shift_rate(J: np.ndarray, w: np.ndarray) -> float:
    """``int dw J(w)/w`` in 1/ps (times hbar: the polaron shift in meV)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(w > 0, J / w, 0.0)
    return float(np.trapezoid(f, w))


def influence_factors(keys: np.ndarray, eta: np.ndarray, dt: float, shift_rate: float = 0.0):
    """``I[k][c_later, c_earlier]`` for k = 0..K (``I[0]`` only its diagonal is used).
    ``shift_rate``: ``int J/w`` -- when non-zero the polaron shift is subtracted, i.e. the
    system energy of level ``l`` is raised by ``hbar * shift_rate * l^2``
    (``Boson_subtract_polaron_shift true``, ``general_system.py:175``)."""
    lp, lm = keys[:, 0], keys[:, 1]
    d = lp - lm
    out = []
    for k, e in enumerate(eta):
        field = e * lp - np.conj(e) * lm                      # depends on the earlier class
        out.append(np.exp(-np.outer(d, field)))
    i0 = np.diag(out[0]).copy()
    if shift_rate:
        i0 = i0 * np.exp(-1j * dt * shift_rate * (lp ** 2 - lm ** 2))
    return out, i0


def _truncate(s: np.ndarray, threshold: float, chi_max: int) -> int:
    keep = int(np.sum(s > threshold * s[0]))
    return max(1, min(keep, chi_max))


def uniform_pt_tensor(I: Sequence[np.ndarray], i0: np.ndarray, threshold: float = 1e-8, chi_max: int = 512,
                      verbose: bool = False) -> np.ndarray:
    """iTEBD contraction of the uniform influence-functional network.

    Two families of lines run through the (time, memory-level) lattice: carrier lines ``A_m``
    (the value of ``c_m`` travelling to later times) and physical lines ``B_n``; at level ``k``
    line ``A_m`` meets ``B_{m+k}`` and picks up ``I_k(c_{m+k}, c_m)``.  Read from the top level
    ``K`` (trivial product state) downwards, every level is one two-site gate
    ``diag(I_k) . SWAP`` on alternating bonds of an infinite MPS with a two-site unit cell; at
    level 0 each pair ``(A_m, B_m)`` is capped with the physical index.  Returns the one-site
    tensor ``f[c, l, r]`` of the uniform MPS  ``F = ... f[c_n] f[c_{n+1}] ...``.
    """
    K = len(I) - 1
    d = I[0].shape[0]
    # Hastings form: B tensors carry their right bond weights; lam[b] is the weight left of B[b]
    B = [np.ones((1, d, 1), dtype=complex) / np.sqrt(d) for _ in range(2)]
    lam = [np.ones(1), np.ones(1)]

    def two_site(x: int, weight: Optional[np.ndarray], swap: bool):
        """gate on the bond between site x (left) and site 1-x (right)."""
        y = 1 - x
        C = np.einsum("lim,mjr->lijr", B[x], B[y])
        if swap:
            C = C.transpose(0, 2, 1, 3)                       # (left, right) <- (right, left)
        if weight is not None:
            # after the swap the left site carries A (earlier), the right site B (later)
            C = C * weight.T[None, :, :, None]                # weight[later, earlier] -> [a, b]
        chi_l, _, _, chi_r = C.shape
        theta = lam[x][:, None, None, None] * C
        U, S, Vh = np.linalg.svd(theta.reshape(chi_l * d, d * chi_r), full_matrices=False)
        keep = _truncate(S, threshold, chi_max)
        S = S[:keep]
        Vh = Vh[:keep]
        nrm = np.linalg.norm(S)
        B[y] = Vh.reshape(keep, d, chi_r)
        B[x] = np.einsum("lijr,kjr->lik", C, np.conj(B[y])) / nrm
        lam[y] = S / nrm
        return keep

    # top level: product state, no swap needed
    bond = 0
    chi = two_site(bond, I[K] if K >= 1 else None, swap=False)
    for k in range(K - 1, 0, -1):
        bond = 1 - bond
        chi = two_site(bond, I[k], swap=True)
        if verbose and (k % 16 == 0 or k == 1):
            print(f"  level {k:4d}: bond dimension {chi}")
    # bring A_m next to B_m and cap the pair with the physical index
    bond = 1 - bond
    two_site(bond, None, swap=True)
    x, y = bond, 1 - bond
    f = np.einsum("lcm,mcr->clr", B[x], B[y]) * i0[:, None, None]
    return f


def resolve_backend(backend: Optional[str] = None) -> str:
    """``"device"`` (iTEBD on the GPU, :mod:`pyaceqd_b200.pt_device`) or ``"host"`` (the NumPy implementation in this
    module).  ``None`` reads ``ACEQD_PT_BUILD`` and defaults to the device whenever a CUDA device is present; asking
    for the device explicitly without one raises."""
    backend = backend or os.environ.get("ACEQD_PT_BUILD", "auto")
    if backend not in ("auto", "host", "device"):
        raise ValueError("PT build backend must be 'auto', 'host' or 'device', not {!r}".format(backend))
    if backend == "auto":
        from . import pt_device
        backend = "device" if pt_device.available() else "host"
    return backend


def uniform_pt(keys: np.ndarray, eta: np.ndarray, dt: float, threshold: float = 1e-8, chi_max: int = 512,
               shift_rate: float = 0.0, verbose: bool = False, backend: str = "host") -> ProcessTensor:
    """Gauge-fixed uniform PT for coupling classes ``keys[n_cls, 2]`` (must contain (0, 0))."""
    keys = np.asarray(keys, dtype=float).reshape(-1, 2)
    null = np.where((np.abs(keys[:, 0]) < 1e-14) & (np.abs(keys[:, 1]) < 1e-14))[0]
    if len(null) == 0:
        raise ValueError("the coupling operator needs an uncoupled level (eigenvalue 0)")
    I, i0 = influence_factors(keys, eta, dt, shift_rate)
    stats = None
    if backend == "device":
        from . import pt_device
        f, stats = pt_device.uniform_pt_tensor(I, i0, threshold, chi_max, verbose=verbose)
    else:
        f = uniform_pt_tensor(I, i0, threshold, chi_max, verbose)
    # boundaries: times before the start / after the end sit in the uncoupled class (all I = 1)
    T0 = f[null[0]]
    ev, vr = np.linalg.eig(T0)
    j = int(np.argmax(np.abs(ev)))
    mu = ev[j]
    f = f / mu
    v_r = vr[:, j]
    evl, vl = np.linalg.eig(T0.T)
    v_l = vl[:, int(np.argmax(np.abs(evl)))]
    v_l = v_l / (v_l @ v_r)
    # gauge: rotate the bond basis so that the left boundary is e_0
    nl = np.linalg.norm(v_l)
    M = np.eye(len(v_l), dtype=complex)
    M[:, 0] = np.conj(v_l) / nl
    Q, _ = np.linalg.qr(M)
    phase = (v_l / nl) @ Q[:, 0]
    Q[:, 0] = Q[:, 0] / phase                                  # now (v_l/nl) @ Q = e_0
    Qi = np.linalg.inv(Q)
    A = np.matmul(np.matmul(Qi[None, :, :], f), Q[None, :, :])  # G f G^-1 with G = Q^-1 (per class: two matrix products)
    q = nl * (Qi @ v_r)
    meta = {"kind": "uniform-itebd", "threshold": threshold, "chi": A.shape[1], "backend": backend}
    if stats is not None:
        meta["device_build"] = stats
    return ProcessTensor(slices=[A], closures=[q], n_initial=0, dt=dt, keys=keys, meta=meta)


def build_gaussian_pt(coupling_diag: Sequence[float], J: np.ndarray, w: np.ndarray, dt: float, t_mem: float,
                      temperature: float, threshold: float = 1e-8, subtract_polaron_shift: bool = True,
                      chi_max: int = 512, dict_zero: float = 1e-12, verbose: bool = False,
                      backend: Optional[str] = None) -> ProcessTensor:
    """PT of a bath with tabulated spectral density ``J(w)`` (1/ps on the grid ``w`` in 1/ps).  ``backend``: see
    :func:`resolve_backend` -- on a GPU box the influence coefficients and the whole iTEBD contraction run on the
    device."""
    _, keys = coupling_classes(np.asarray(coupling_diag, dtype=float), dict_zero)
    K = max(1, int(round(t_mem / dt)))
    backend = resolve_backend(backend)
    if backend == "device":
        from . import pt_device
        eta = pt_device.eta_coefficients(J, w, dt, K, temperature)
    else:
        eta = eta_coefficients(J, w, dt, K, temperature)
    shift = polaron_shift_rate(J, w) if subtract_polaron_shift else 0.0
    pt = uniform_pt(keys, eta, dt, threshold=threshold, chi_max=chi_max, shift_rate=shift, verbose=verbose,
                    backend=backend)
    pt.meta["coupling_diag"] = np.asarray(coupling_diag, dtype=float)     # travels with the file (ace_cli)
    return pt


def build_qd_phonon_pt(coupling_diag: Sequence[float], dt: float, t_mem: float, a_e: float = 5.0,
                       a_h: Optional[float] = None, temperature: float = 4.0, threshold: float = 1e-8,
                       e_max: float = 7.0, use_infinite: bool = True, chi_max: int = 512, n_w: int = 20001,
                       verbose: bool = False, backend: Optional[str] = None) -> ProcessTensor:
    """GaAs QD / LA-phonon PT with the parameters the reference passes to ACE
    (``general_system.py:159-192``): ``Boson_E_min 0``, ``Boson_E_max e_max`` (meV)."""
    w = np.linspace(0.0, e_max / constants.hbar, n_w)
    J = qd_phonon_spectral_density(w, a_e, a_h)
    pt = build_gaussian_pt(coupling_diag, J, w, dt, t_mem, temperature, threshold=threshold,
                           subtract_polaron_shift=True, chi_max=chi_max, verbose=verbose, backend=backend)
    pt.meta.update({"a_e": a_e, "a_h": a_h, "temperature": temperature, "t_mem": t_mem,
                    "use_infinite": bool(use_infinite)})
    if verbose:
        print(f"QD phonon PT: dt={dt} ps, memory {t_mem} ps, T={temperature} K, chi={pt.chi_max}")
    return pt


# ---------------------------------------------------------------------------------------------- spectral-density files
def write_spectral_density(path: str, a_e: float = 5.0, a_h: Optional[float] = None, e_min: float = 0.0,
                           e_max: float = 15.0, n: int = 2000) -> None:
    """``Boson_J_print <file> 0 15 2000`` (reference ``general_system.py:186-187``): two columns, ``hbar w`` in meV and
    ``J(w)`` in 1/ps, ``n`` rows.  (ACE's own print format cannot be checked here; this is the format
    :func:`read_spectral_density` reads back.)"""
    e = np.linspace(e_min, e_max, n)
    J = qd_phonon_spectral_density(e / constants.hbar, a_e, a_h)
    np.savetxt(path, np.column_stack([e, J]), fmt="%.12e", delimiter=" ")


def read_spectral_density(path: str):
    """Tabulated spectral density (``Boson_J_from_file``, reference ``general_system.py:178-179``): columns ``hbar w``
    (meV) and ``J`` (1/ps).  Returns ``(w [1/ps], J [1/ps])`` on the file's grid."""
    data = np.loadtxt(path, ndmin=2)
    if data.shape[1] < 2 or len(data) < 2:
        raise ValueError("{}: expected two columns (energy in meV, J in 1/ps)".format(path))
    order = np.argsort(data[:, 0])
    return data[order, 0] / constants.hbar, data[order, 1]


def build_pt_from_spectral_density_file(path: str, coupling_diag: Sequence[float], dt: float, t_mem: float,
                                        temperature: float, threshold: float = 1e-8, e_max: Optional[float] = None,
                                        n_w: int = 20001, chi_max: int = 512, verbose: bool = False,
                                        backend: Optional[str] = None) -> ProcessTensor:
    """PT of the bath whose spectral density is tabulated in ``path``; the table is interpolated linearly onto the
    builder's frequency grid and cut at ``e_max`` (meV, ``Boson_E_max``) or at the end of the table."""
    w_tab, j_tab = read_spectral_density(path)
    w_hi = w_tab[-1] if e_max is None else min(w_tab[-1], e_max / constants.hbar)
    w = np.linspace(0.0, w_hi, n_w)
    J = np.interp(w, w_tab, j_tab, left=0.0, right=0.0)
    pt = build_gaussian_pt(coupling_diag, J, w, dt, t_mem, temperature, threshold=threshold, subtract_polaron_shift=True,
                           chi_max=chi_max, verbose=verbose, backend=backend)
    pt.meta.update({"J_file": os.path.basename(path), "temperature": temperature, "t_mem": t_mem})
    return pt

