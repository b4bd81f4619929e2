"""ctypes wrapper of the C oracle (oracle_c.c).  Test infrastructure / CPU baseline only --
see the header of oracle_c.c.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_c_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_c_max_threads())


def expm(mats: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(mats, dtype=np.complex128)
    if a.ndim == 2:
        a = a[None]
    out = np.empty_like(a)
    rc = lib().oracle_c_expm(ctypes.c_int(a.shape[1]), ctypes.c_int(a.shape[0]),
                             ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(out.ctypes.data))
    if rc:
        raise RuntimeError(f"oracle_c_expm failed ({rc})")
    return out


def propagate_sweep(problem, pt, jobs, t_eval="half_mid", n_threads=0):
    """MTO-free trajectories starting at the PT origin (pulse sweeps).
    Returns a list of ``[n_out, n_steps+1]`` complex arrays."""
    offs = {"half_mid": (0.25, 0.75), "step_mid": (0.5, 0.5), "start": (0.0, 0.5)}[t_eval]
    if any(j.mtos for j in jobs):
        raise ValueError("the C oracle handles MTO-free trajectories only (use oracle.propagate)")
    c128 = lambda a: np.ascontiguousarray(a, dtype=np.complex128)
    NL, n_out = problem.L0.shape[0], problem.out_w.shape[0]
    blk = np.ascontiguousarray(pt.block_of_class(problem.cls_keys)[np.asarray(problem.cls)], dtype=np.int32)
    tab_index = {"x": 0, "y": 1, "rf": 2}
    ft = np.asarray([tab_index[p] for p in problem.field_pol], dtype=np.int32)
    nmax = max([len(tb.values) for j in jobs for tb in j.tables.values()] + [1])
    grid = None
    packed = np.zeros((len(jobs), 3, nmax), dtype=np.complex128)
    for s, j in enumerate(jobs):
        for pol, tb in j.tables.items():
            if grid is None:
                grid = (tb.t0, tb.dt)
            n = len(tb.values)
            packed[s, tab_index[pol], :n] = tb.values
            packed[s, tab_index[pol], n:] = tb.values[-1] if n else 0
    if grid is None:
        grid = (0.0, 1.0)
    n_steps = np.asarray([j.n_steps for j in jobs], dtype=np.int32)
    rows = n_steps.astype(np.int64) + 1
    out_off = np.zeros(len(jobs), dtype=np.int64)
    out_off[1:] = np.cumsum(rows[:-1] * n_out)
    out = np.zeros(int(np.sum(rows * n_out)), dtype=np.complex128)
    sl = [c128(s) for s in pt.slices]
    cl = [c128(q) for q in pt.closures]
    ns = len(sl)
    sp = (ctypes.c_void_p * ns)(*[s.ctypes.data for s in sl])
    cp = (ctypes.c_void_p * ns)(*[q.ctypes.data for q in cl])
    chi_in = np.asarray([s.shape[1] for s in sl], dtype=np.int32)
    chi_out = np.asarray([s.shape[2] for s in sl], dtype=np.int32)
    L0, LA, LB, ow, r0 = c128(problem.L0), c128(problem.LA), c128(problem.LB), c128(problem.out_w), c128(problem.rho0)
    sets = np.arange(len(jobs), dtype=np.int32)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    I, D = ctypes.c_int, ctypes.c_double
    rc = lib().oracle_c_propagate(
        I(NL), I(len(ft)), I(n_out), P(L0), P(LA), P(LB), P(ft), P(ow), P(blk), I(sl[0].shape[0]), I(ns),
        I(pt.n_initial), P(chi_in), P(chi_out), sp, cp, I(len(jobs)), P(sets), P(n_steps), D(jobs[0].dt),
        D(jobs[0].t_start), I(len(jobs)), I(3), I(nmax), D(grid[0]), D(grid[1]), P(packed), P(r0),
        D(offs[0]), D(offs[1]), P(out_off), P(out), I(n_threads))
    if rc:
        raise RuntimeError(f"oracle_c_propagate failed ({rc})")
    return [np.ascontiguousarray(out[out_off[i]: out_off[i] + rows[i] * n_out].reshape(rows[i], n_out).T)
            for i in range(len(jobs))]
