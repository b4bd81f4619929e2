"""ctypes binding of the on-device process-tensor builder (``csrc/ptbuild.cu`` -> ``libaceqd_ptbuild.so``,
C ABI ``include/aceqd_ptbuild.h``).

The GPU replacement of the PT-generation run the reference delegates to ACE
(``pyaceqd/general_system/general_system.py:159-192``); :mod:`pyaceqd_b200.pt_builder` holds the NumPy implementation of
the same algorithm, the spectral densities and the gauge fixing of the result, and calls into this module when a CUDA
device is present (``ACEQD_PT_BUILD=host|device`` overrides)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_void_p
from typing import Sequence

import numpy as np

from . import constants

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libaceqd_ptbuild.so")
_lib = None


class PtBuildError(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtBuildError(f"{LIB_PATH} not found: build it with `python -m pyaceqd_b200.build`")
    lib = ctypes.CDLL(LIB_PATH)
    lib.aceqd_ptbuild_last_error.restype = c_char_p
    lib.aceqd_ptbuild_device_count.restype = c_int
    lib.aceqd_ptbuild_eta.argtypes = [c_int, c_int, c_void_p, c_void_p, c_double, c_int, c_double, c_void_p]
    lib.aceqd_ptbuild_uniform.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_int, c_int, c_void_p,
                                          POINTER(c_int), c_void_p]
    _lib = lib
    return lib


def available() -> bool:
    """True when the library loads and a CUDA device is visible."""
    try:
        return load_library().aceqd_ptbuild_device_count() > 0
    except (OSError, PtBuildError):
        return False


def _check(rc: int, what: str):
    if rc != 0:
        raise PtBuildError("{} failed ({}): {}".format(what, rc, load_library().aceqd_ptbuild_last_error().decode()))


def _device() -> int:
    return int(os.environ.get("ACEQD_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def eta_coefficients(J: np.ndarray, w: np.ndarray, dt: float, K: int, temperature: float) -> np.ndarray:
    """``pt_builder.eta_coefficients`` on the device (``k_eta``: one CTA per memory step)."""
    lib = load_library()
    if lib.aceqd_ptbuild_device_count() <= 0:
        raise PtBuildError("the device PT builder needs a CUDA device (ACEQD_PT_BUILD=host selects the NumPy builder)")
    w = np.ascontiguousarray(w, dtype=np.float64)
    J = np.ascontiguousarray(J, dtype=np.float64)
    eta = np.zeros(K + 1, dtype=np.complex128)
    h2kt = constants.hbar / (2 * constants.kB * temperature) if temperature > 0 else 0.0
    _check(lib.aceqd_ptbuild_eta(_device(), len(w), w.ctypes.data, J.ctypes.data, float(dt), int(K), float(h2kt),
                                 eta.ctypes.data), "aceqd_ptbuild_eta")
    return eta


def uniform_pt_tensor(I: Sequence[np.ndarray], i0: np.ndarray, threshold: float = 1e-8, chi_max: int = 512,
                      svd_method: int = None, verbose: bool = False):
    """``pt_builder.uniform_pt_tensor`` on the device.  Returns ``(f[c, l, r], stats)``."""
    lib = load_library()
    if lib.aceqd_ptbuild_device_count() <= 0:
        raise PtBuildError("the device PT builder needs a CUDA device (ACEQD_PT_BUILD=host selects the NumPy builder)")
    if svd_method is None:
        svd_method = int(os.environ.get("ACEQD_PT_SVD", "0"))
    K = len(I) - 1
    d = I[0].shape[0]
    weights = np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.complex128) for m in I]))      # [K+1][later][earlier]
    i0 = np.ascontiguousarray(i0, dtype=np.complex128)
    f = np.zeros((d, chi_max, chi_max), dtype=np.complex128)
    chi = c_int(0)
    stats = np.zeros(4)
    _check(lib.aceqd_ptbuild_uniform(_device(), d, K, weights.ctypes.data, i0.ctypes.data, float(threshold), int(chi_max),
                                     int(svd_method), f.ctypes.data, ctypes.byref(chi), stats.ctypes.data),
           "aceqd_ptbuild_uniform")
    info = {"total_ms": float(stats[0]), "svd_ms": float(stats[1]), "gemm_ms": float(stats[2]),
            "largest_svd_dim": int(stats[3]), "svd": "gesvd" if svd_method == 0 else "gesvdp", "levels": K}
    if verbose:
        print("  device iTEBD: chi = {}, {:.1f} ms ({:.1f} ms in {} SVDs up to {} rows)".format(
            chi.value, info["total_ms"], info["svd_ms"], K + 1, info["largest_svd_dim"]))
    return np.ascontiguousarray(f[:, :chi.value, :chi.value]), info
