"""ctypes binding of libaceqd.so and the host-side batch planner.

The planner turns a list of :class:`~pyaceqd_b200.jobs.Job` (one per ``system(...)`` call of the
reference, e.g. the fan-out of ``pyaceqd/two_time/correlations.py:153-170``) into one
``aceqd_batch``: drive tables, operator sequences, explicit MTO entries, trajectory descriptors
and the tiling of trajectories onto persistent CTAs.  Jobs that share drive tables and start
time are forked from one common trunk: the trunk is propagated once, the bond state
is snapshotted at each job's first multi-time-operator step, and every job continues from its
snapshot -- exact, because the full system x bond state is carried.

There is no CPU fallback: importing works anywhere, but every compute call needs the CUDA
library and a B200.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_longlong, c_void_p
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import planner
from .jobs import FieldTable, Job
from .problem import Problem
from .process_tensor import ProcessTensor, trivial_pt

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libaceqd.so")
MAX_OVR = 6
N_SM = 148

# step kernels (aceqd_batch.kernel): "dmma" = the library's choice between the persistent tile kernel and the
# small-bond kernel (and, for large Liouville spaces, the planner's choice of the split-K cluster kernel), "check" =
# plain-FMA check kernel, "splitk" = split-K cluster kernel, "small" / "tile" force k_step_small / k_step_dmma
KERNELS = {"dmma": 0, "check": 1, "splitk": 3, "small": 4, "tile": 5}
SMALL_MIN_TRAJ = 2048     # ACEQD_SMALL_MIN_TRAJ of include/aceqd.h


def resolve_kernel(kernel: str, NL: int) -> str:
    """``"auto"`` (or the ``ACEQD_KERNEL`` environment override) -> ``"dmma"``: the persistent kernels, with the
    planner (:meth:`Engine._tile_and_cluster`) choosing tile size, cluster size and the column-split variant."""
    if kernel == "auto":
        kernel = os.environ.get("ACEQD_KERNEL", "auto")
    if kernel == "auto":
        return "dmma"
    if kernel not in KERNELS:
        raise ValueError(f"unknown kernel {kernel!r}; choose from {sorted(KERNELS)} or 'auto'")
    return kernel

T_EVAL = {"half_mid": (0.25, 0.75), "step_mid": (0.5, 0.5), "start": (0.0, 0.5)}

SEQ_DT = np.dtype([("set", "<i4"), ("step0", "<i4"), ("len", "<i4"), ("first_has_prev", "<i4")], align=True)
ENTRY_DT = np.dtype([("set", "<i4"), ("step", "<i4"), ("sb", "<i4"), ("sa", "<i4"), ("has_prev", "<i4"),
                     ("clamp", "<i4")], align=True)
TLSEG_DT = np.dtype([("start", "<i4"), ("count", "<i4"), ("emit", "<i4"), ("stride", "<i4")], align=True)
TRAJ_DT = np.dtype([("ent0", "<i8"), ("out_off", "<i8"), ("step0", "<i4"), ("n_steps", "<i4"),
                    ("init_kind", "<i4"), ("init_index", "<i4"), ("n_ovr", "<i4"),
                    ("ovr_step", "<i4", (MAX_OVR,)), ("ovr_ent", "<i4", (MAX_OVR,)),
                    ("snap_off", "<i4"), ("snap_cnt", "<i4"), ("snap_slot0", "<i4"), ("out_from", "<i4")],
                   align=True)


class _Batch(ctypes.Structure):
    _fields_ = [
        ("dt", c_double), ("t0", c_double), ("eval_off1", c_double), ("eval_off2", c_double),
        ("n_sets", c_int32), ("n_tables", c_int32), ("n_samples", c_int32),
        ("tab_t0", c_double), ("tab_dt", c_double), ("tables", c_void_p),
        ("n_seq", c_int32), ("seqs", c_void_p),
        ("n_entries", c_int32), ("entries", c_void_p),
        ("n_mto_mats", c_int32), ("mto_mats", c_void_p),
        ("n_rho0", c_int32), ("rho0s", c_void_p),
        ("n_traj", c_int32), ("trajs", c_void_p),
        ("tile_T", c_int32), ("n_tiles", c_int32), ("tile_traj", c_void_p),
        ("n_snap_steps", c_int32), ("snap_steps", c_void_p), ("n_snap_slots", c_int32),
        ("out_elems", c_int64), ("out", c_void_p),
        ("device_resident", c_int32), ("kernel", c_int32), ("cluster", c_int32), ("n_reduce", c_int32),
        ("reduce_ch", c_void_p), ("reduce_spacing", c_double), ("reduce_out", c_void_p),
    ]


class EngineError(RuntimeError):
    pass


_lib_handle = None


def load_library():
    """Load the in-tree CUDA library; fail loudly if it is missing (no fallback)."""
    global _lib_handle
    if _lib_handle is not None:
        return _lib_handle
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} not found: build it with `python -m pyaceqd_b200.build` "
            "(the engine has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.aceqd_last_error.restype = c_char_p
    lib.aceqd_version.restype = c_char_p
    lib.aceqd_ctx_create.argtypes = [c_int, c_void_p, POINTER(c_void_p)]
    lib.aceqd_ctx_destroy.argtypes = [c_void_p]
    lib.aceqd_ctx_destroy.restype = None
    lib.aceqd_ctx_sync.argtypes = [c_void_p]
    lib.aceqd_launch_count.argtypes = [c_void_p]
    lib.aceqd_launch_count.restype = c_longlong
    for name in ("aceqd_last_step_kernel", "aceqd_last_opbuild_kernel", "aceqd_last_other_kernel"):
        getattr(lib, name).argtypes = [c_void_p]
        getattr(lib, name).restype = c_char_p
    lib.aceqd_last_timings.argtypes = [c_void_p, POINTER(c_float), POINTER(c_float)]
    lib.aceqd_pt_create.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, POINTER(c_void_p)]
    lib.aceqd_pt_destroy.argtypes = [c_void_p]
    lib.aceqd_pt_destroy.restype = None
    lib.aceqd_pt_chi_pad.argtypes = [c_void_p]
    lib.aceqd_problem_create.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.aceqd_problem_destroy.argtypes = [c_void_p]
    lib.aceqd_problem_destroy.restype = None
    for name in ("aceqd_propagate_batch", "aceqd_run_steps"):
        getattr(lib, name).argtypes = [c_void_p, c_void_p, c_void_p, POINTER(_Batch)]
    lib.aceqd_build_operators.argtypes = [c_void_p, c_void_p, POINTER(_Batch)]
    lib.aceqd_snapshot_read.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p]
    lib.aceqd_expm_batch.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.aceqd_tlmap_run.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.aceqd_tlmap_last_ms.argtypes = [c_void_p, POINTER(c_float)]
    lib.aceqd_max_tile.argtypes = [c_int, c_int]
    lib.aceqd_max_tile_global_pt.argtypes = [c_int, c_int]
    lib.aceqd_splitk_fit.argtypes = [c_int, c_int, c_int, c_int]
    lib.aceqd_splitk_fit.restype = c_longlong
    lib.aceqd_pass_load.argtypes = [c_void_p, c_int, c_int]
    lib.aceqd_fp64_peak.argtypes = [c_void_p, c_int, c_int, POINTER(c_double)]
    lib.aceqd_host_alloc.argtypes = [ctypes.c_size_t, POINTER(c_void_p)]
    lib.aceqd_host_free.argtypes = [c_void_p]
    lib.aceqd_host_free.restype = None
    lib.aceqd_struct_sizes.argtypes = [POINTER(c_int32)]
    lib.aceqd_struct_sizes.restype = None
    sizes = (c_int32 * 4)()
    lib.aceqd_struct_sizes(sizes)
    want = (SEQ_DT.itemsize, ENTRY_DT.itemsize, TRAJ_DT.itemsize, ctypes.sizeof(_Batch))
    if tuple(sizes) != want:
        raise EngineError(f"ABI mismatch between engine.py {want} and libaceqd.so {tuple(sizes)}")
    _lib_handle = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().aceqd_last_error().decode(errors="replace")
        raise EngineError(f"{what} failed (status {rc}): {msg}")


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


def tail_trapezoid(block: np.ndarray, pairs, spacing: float) -> np.ndarray:
    """Host statement of the fused tail reduction: ``block[n_out, rows]`` -> ``[len(pairs)]``."""
    out = np.zeros(len(pairs), dtype=np.complex128)
    rows = block.shape[1]
    if rows < 2:
        return out
    w = np.ones(rows)
    w[0] = w[-1] = 0.5
    for p, (ch_tau, ch_zero) in enumerate(pairs):
        col = np.array(block[ch_tau], dtype=complex)
        col[0] = block[ch_zero][0]
        out[p] = spacing * np.sum(w * col)
    return out


def class_sorted_positions(block_of_alpha: np.ndarray) -> np.ndarray:
    """Position of every Liouville index in PT-block-sorted order (stable)."""
    order = np.argsort(block_of_alpha, kind="stable")
    pos = np.empty_like(order)
    pos[order] = np.arange(len(order))
    return pos.astype(np.int32)


def choose_tile(n_traj: int, rows_per_block: Sequence[int], t_max: int, n_sm: int = N_SM) -> int:
    """Trajectories per persistent CTA: minimise waves x DMMA m-tiles per tile."""
    best_t, best_cost = 1, None
    t = 1
    while t <= t_max:
        tiles = -(-n_traj // t)
        waves = -(-tiles // n_sm)
        mtiles = sum(-(-(t * r) // 8) for r in rows_per_block)
        cost = waves * mtiles
        if best_cost is None or cost <= best_cost:  # ties -> larger tile (less L2 traffic)
            best_t, best_cost = t, cost
        t *= 2
    return best_t


CLUSTER_CAPACITY = {1: N_SM, 2: N_SM, 4: 132, 8: 128, 16: 64}   # co-resident CTAs per cluster size (GPC packing)
SPLITK_PASS_OVERHEAD = 128.0   # split-K planner: per-pass exchange cost in units of (m-tile x PT k-row)


GLOBAL_PT_PENALTY = 1.15        # a tile without a shared-memory PT ring (fragments from L2) against one with it


# 16-CTA clusters (beyond the portable limit of 8; a whole GPC) serve the lone trunk of a G2 map: 9 coupling classes on
# 16 SMs = one GEMM pass per CTA and step (cfg3 trunk 4.3 -> 3.3 ms).  ACEQD_CLUSTER16=0 keeps clusters <= 8.
DEFAULT_CLUSTERS = (1, 2, 4, 8) if os.environ.get("ACEQD_CLUSTER16") == "0" else (1, 2, 4, 8, 16)


def choose_tile_cluster(n_traj: int, pass_load, t_max: int, clusters: Sequence[int] = DEFAULT_CLUSTERS,
                        t_ring: Optional[int] = None) -> Tuple[int, int]:
    """(trajectories per tile, CTAs per tile): minimise waves x (m-tiles of the most loaded CTA per
    step).  ``pass_load(T, C)`` is ``aceqd_pass_load``.  A cluster splits a tile's GEMM passes over C
    SMs, which only pays when the batch alone cannot fill the GPU; it costs a row exchange per step and
    repeats the small system product in every CTA (25 % per doubling assumed; measured:
    profiles/r01k_*), and ties go to the smaller cluster and then to the larger tile."""
    best, best_cost = (1, 1), None
    t = 1
    while t <= t_max:
        tiles = -(-n_traj // t)
        for c in clusters:
            load = pass_load(t, c)
            if load <= 0:
                continue
            if t > max(1, n_traj):
                continue     # no point in tiles wider than the batch
            waves = -(-(tiles * c) // CLUSTER_CAPACITY[c])
            cost = waves * load * (1.0 + 0.25 * (c.bit_length() - 1))
            if t_ring is not None and t > t_ring:      # tiles this large leave no room for the PT chunk ring
                cost *= GLOBAL_PT_PENALTY
            if best_cost is None or cost < best_cost - 1e-9 or (abs(cost - best_cost) <= 1e-9 and c <= best[1]):
                best, best_cost = (t, c), cost
        t *= 2
    return best


@dataclass
class _Plan:
    """Host arrays of one aceqd_batch (kept alive for the duration of the call)."""
    batch: _Batch
    keep: list
    out: np.ndarray
    out_off: np.ndarray
    n_rows: np.ndarray


class Engine:
    """One CUDA context (device + stream + workspace) with PT / problem handle caches."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        h = c_void_p()
        _check(self.lib.aceqd_ctx_create(int(device), c_void_p(stream) if stream else None,
                                         ctypes.byref(h)), "aceqd_ctx_create")
        self.ctx = h
        self.device = int(device)
        self._pinned: list = []
        self._out_pinned = None
        self._pts: Dict[int, Tuple[c_void_p, ProcessTensor]] = {}
        self._probs: Dict[Tuple[int, int], Tuple[c_void_p, Problem, np.ndarray]] = {}
        self.record_timings = False      # bench: log device times of every run_jobs launch
        self.timing_log: List[dict] = []

    # -------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "ctx", None):
            for h, _ in self._pts.values():
                self.lib.aceqd_pt_destroy(h)
            for h, _, _ in self._probs.values():
                self.lib.aceqd_problem_destroy(h)
            self._pts.clear()
            self._probs.clear()
            self._out_pinned = None
            for ptr in self._pinned:
                self.lib.aceqd_host_free(ptr)
            self._pinned = []
            self.lib.aceqd_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------- handles
    def pt_handle(self, pt: ProcessTensor) -> c_void_p:
        key = id(pt)
        if key in self._pts:
            return self._pts[key][0]
        ns = pt.n_slices
        chi_in = np.asarray([s.shape[1] for s in pt.slices], dtype=np.int32)
        chi_out = np.asarray([s.shape[2] for s in pt.slices], dtype=np.int32)
        sl = [_c128(s) for s in pt.slices]
        cl = [_c128(q) for q in pt.closures]
        sp = (c_void_p * ns)(*[s.ctypes.data for s in sl])
        cp = (c_void_p * ns)(*[q.ctypes.data for q in cl])
        h = c_void_p()
        _check(self.lib.aceqd_pt_create(self.ctx, pt.n_cls, ns, pt.n_initial, chi_in.ctypes.data,
                                        chi_out.ctypes.data, sp, cp, ctypes.byref(h)), "aceqd_pt_create")
        self._pts[key] = (h, pt)
        return h

    def problem_handle(self, prob: Problem, pt: ProcessTensor) -> Tuple[c_void_p, np.ndarray]:
        key = (id(prob), id(pt))
        if key in self._probs:
            return self._probs[key][0], self._probs[key][2]
        blk_of_cls = pt.block_of_class(prob.cls_keys)
        blk_of_alpha = np.ascontiguousarray(blk_of_cls[np.asarray(prob.cls)], dtype=np.int32)
        pos = np.ascontiguousarray(class_sorted_positions(blk_of_alpha), dtype=np.int32)
        tab_index = {"x": 0, "y": 1, "rf": 2}
        ft = np.asarray([tab_index[p] for p in prob.field_pol], dtype=np.int32)
        L0, LA, LB, ow = _c128(prob.L0), _c128(prob.LA), _c128(prob.LB), _c128(prob.out_w)
        h = c_void_p()
        _check(self.lib.aceqd_problem_create(
            self.ctx, prob.NL, prob.n_fields, prob.n_out, L0.ctypes.data,
            LA.ctypes.data if prob.n_fields else None, LB.ctypes.data if prob.n_fields else None,
            ft.ctypes.data if prob.n_fields else None, ow.ctypes.data if prob.n_out else None,
            pos.ctypes.data, blk_of_alpha.ctypes.data, ctypes.byref(h)), "aceqd_problem_create")
        self._probs[key] = (h, prob, blk_of_alpha)
        return h, blk_of_alpha

    # -------------------------------------------------------------- small services
    def expm(self, mats: np.ndarray) -> np.ndarray:
        """Batched matrix exponential on the device (the operator builder's kernel)."""
        a = _c128(mats)
        if a.ndim == 2:
            a = a[None]
        out = np.empty_like(a)
        _check(self.lib.aceqd_expm_batch(self.ctx, a.shape[1], a.shape[0], a.ctypes.data,
                                         out.ctypes.data), "aceqd_expm_batch")
        return out

    def tlmap_run(self, mats: np.ndarray, v0: np.ndarray, seg_off: np.ndarray, segs: np.ndarray,
                  w: Optional[np.ndarray] = None, n_emit_max: int = 0, want_final: bool = False):
        """Batched time-local map chains (``aceqd_tlmap_run``): ``mats[n_mats, NL, NL]`` matrix pool,
        ``v0[n_chains, NL]`` start vectors, chain ``i`` runs segments ``segs[seg_off[i]:seg_off[i+1]]``
        (dtype ``TLSEG_DT``).  Returns ``(out[n_chains, n_emit_max, n_w] | None, final[n_chains, NL] | None)``."""
        mats, v0 = _c128(mats), _c128(v0)
        NL = mats.shape[-1]
        n_chains = v0.shape[0]
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        segs = np.ascontiguousarray(segs, dtype=TLSEG_DT)
        w = _c128(w).reshape(-1, NL) if w is not None else np.zeros((0, NL), complex)
        out = np.zeros((n_chains, n_emit_max, w.shape[0]), dtype=np.complex128) if n_emit_max and len(w) else None
        fin = np.empty((n_chains, NL), dtype=np.complex128) if want_final else None
        if out is None and fin is None:
            raise ValueError("nothing to compute: ask for emitted outputs and/or final vectors")
        _check(self.lib.aceqd_tlmap_run(self.ctx, NL, mats.shape[0], mats.ctypes.data, n_chains, v0.ctypes.data,
                                        seg_off.ctypes.data, len(segs), segs.ctypes.data if len(segs) else None,
                                        w.shape[0], w.ctypes.data if len(w) else None, int(n_emit_max),
                                        out.ctypes.data if out is not None else None,
                                        fin.ctypes.data if fin is not None else None), "aceqd_tlmap_run")
        return out, fin

    def tlmap_last_ms(self) -> float:
        v = c_float()
        _check(self.lib.aceqd_tlmap_last_ms(self.ctx, ctypes.byref(v)), "aceqd_tlmap_last_ms")
        return v.value

    def fp64_peak(self, kind: str = "dmma", iters: int = 20000) -> float:
        v = c_double()
        # "dmma", "dfma", or an int: 10 + w = DMMA with ONE block per SM and w warps per SM sub-partition
        code = kind if isinstance(kind, int) else (0 if kind == "dmma" else 1)
        _check(self.lib.aceqd_fp64_peak(self.ctx, code, iters, ctypes.byref(v)),
               "aceqd_fp64_peak")
        return v.value

    def launch_count(self) -> int:
        return int(self.lib.aceqd_launch_count(self.ctx))

    def last_kernels(self) -> Dict[str, str]:
        """Names of the kernel instantiations the last step / operator-builder / service launch used."""
        return {k: getattr(self.lib, "aceqd_last_%s_kernel" % k)(self.ctx).decode() for k in ("step", "opbuild", "other")}

    def last_timings(self) -> Tuple[float, float]:
        a, b = c_float(), c_float()
        _check(self.lib.aceqd_last_timings(self.ctx, ctypes.byref(a), ctypes.byref(b)), "aceqd_last_timings")
        return a.value, b.value

    def max_tile(self, NL: int, chi_pad: int) -> int:
        """Largest tile the step kernel can hold, with or without a shared-memory ring for the PT chunks."""
        return max(int(self.lib.aceqd_max_tile(NL, chi_pad)), int(self.lib.aceqd_max_tile_global_pt(NL, chi_pad)))

    def max_tile_ring(self, NL: int, chi_pad: int) -> int:
        return int(self.lib.aceqd_max_tile(NL, chi_pad))

    def read_snapshot(self, slot: int, NL: int, chi_pad: int) -> np.ndarray:
        out = np.empty((NL, chi_pad), dtype=np.complex128)
        _check(self.lib.aceqd_snapshot_read(self.ctx, slot, NL, chi_pad, out.ctypes.data), "aceqd_snapshot_read")
        return out

    # -------------------------------------------------------------- pinned host memory
    def pinned_empty(self, shape, dtype=np.complex128) -> np.ndarray:
        """Page-locked host array (owned by the engine; freed with it)."""
        dt_ = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        ptr = c_void_p()
        _check(self.lib.aceqd_host_alloc(max(1, n) * dt_.itemsize, ctypes.byref(ptr)), "aceqd_host_alloc")
        self._pinned.append(ptr)
        buf = (ctypes.c_char * (max(1, n) * dt_.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt_, count=n).reshape(shape)

    def _pinned_out(self, n_elems: int) -> np.ndarray:
        if self._out_pinned is None or self._out_pinned.size < n_elems:
            self._out_pinned = self.pinned_empty(int(n_elems * 1.1) + 16, np.complex128)
        return self._out_pinned[:n_elems]

    # -------------------------------------------------------------- uniform sweeps (vectorised)
    def plan_sweep(self, prob: Problem, pt: ProcessTensor, n_traj: int, n_steps: int, dt: float,
                   t_start: float, n_sets: int, n_samples: int, grid: Tuple[float, float], *,
                   sets: Optional[np.ndarray] = None, kernel: str = "auto", t_eval: str = "half_mid",
                   tile_T: Optional[int] = None, cluster: Optional[int] = None) -> "_Plan":
        """Descriptors of a pulse-parameter sweep: `n_traj` MTO-free trajectories of equal length
        starting at the PT origin (reference fan-out
        ``two_level_system/rabi_rotations.py:172-198``).  Table / output pointers are filled in
        by :meth:`run_sweep`."""
        NL, n_out = prob.NL, prob.n_out
        kernel = resolve_kernel(kernel, NL)
        _, blk_of_alpha = self.problem_handle(prob, pt)
        chi_pad = -(-pt.chi_max // 8) * 8
        t_max = self.max_tile(NL, chi_pad)
        if t_max < 1 and kernel in ("dmma", "tile"):
            kernel = "splitk"     # one trajectory does not fit a CTA: spread its bond columns over a cluster
        T, C = self._tile_and_cluster(prob, pt, n_traj, t_max, tile_T, cluster, kernel)
        n_tiles = -(-n_traj // T)
        sets = np.arange(n_traj, dtype=np.int32) if sets is None else np.asarray(sets, dtype=np.int32)
        seqs = np.zeros(n_traj, dtype=SEQ_DT)
        seqs["set"], seqs["step0"], seqs["len"] = sets, 0, n_steps + 1
        trajs = np.zeros(n_traj, dtype=TRAJ_DT)
        trajs["ent0"] = np.arange(n_traj, dtype=np.int64) * (n_steps + 1)
        trajs["out_off"] = np.arange(n_traj, dtype=np.int64) * (n_steps + 1) * n_out
        trajs["n_steps"] = n_steps
        tile_traj = np.full(n_tiles * T, -1, dtype=np.int32)
        tile_traj[:n_traj] = np.arange(n_traj, dtype=np.int32)
        rho0 = _c128(prob.rho0).reshape(1, NL)
        off1, off2 = T_EVAL[t_eval]
        b = _Batch()
        b.dt, b.t0, b.eval_off1, b.eval_off2 = float(dt), float(t_start), off1, off2
        b.n_sets, b.n_tables, b.n_samples = int(n_sets), 3, int(n_samples)
        b.tab_t0, b.tab_dt = float(grid[0]), float(grid[1])
        b.n_seq, b.seqs = n_traj, seqs.ctypes.data
        b.n_entries, b.entries, b.n_mto_mats, b.mto_mats = 0, None, 0, None
        b.n_rho0, b.rho0s = 1, rho0.ctypes.data
        b.n_traj, b.trajs = n_traj, trajs.ctypes.data
        b.tile_T, b.n_tiles, b.tile_traj = T, n_tiles, tile_traj.ctypes.data
        b.n_snap_steps, b.snap_steps, b.n_snap_slots = 0, None, 0
        b.out_elems = int(n_traj) * (n_steps + 1) * n_out
        b.kernel = KERNELS[kernel]
        b.cluster = C
        return _Plan(batch=b, keep=[seqs, trajs, tile_traj, rho0], out=None,
                     out_off=trajs["out_off"], n_rows=trajs["n_steps"] + 1)

    def _tile_and_cluster(self, prob, pt, n_traj, t_max, tile_T=None, cluster=None, kernel="dmma") -> Tuple[int, int]:
        hp, _ = self.problem_handle(prob, pt)
        if kernel == "small":      # one warp per 8 trajectories: the tile list only orders the octets
            return min(tile_T or 8, t_max), 1
        if kernel == "splitk":
            return self._splitk_tile(prob, pt, n_traj, tile_T, cluster)[:2]
        load = lambda t, c: int(self.lib.aceqd_pass_load(hp, t, c))
        if tile_T and cluster:
            return min(tile_T, t_max), cluster
        if tile_T:
            t = min(tile_T, t_max)
            return t, choose_tile_cluster(n_traj, lambda tt, c: load(t, c) if tt == t else 0, t)[1]
        t_ring = self.max_tile_ring(prob.NL, -(-pt.chi_max // 8) * 8)
        return choose_tile_cluster(n_traj, load, t_max, clusters=(cluster,) if cluster else DEFAULT_CLUSTERS, t_ring=t_ring)

    def _splitk_tile(self, prob, pt, n_traj, tile_T=None, cluster=None):
        """(G, C, cost) of the split-K cluster kernel: G trajectories per tile on a cluster of C CTAs that hold
        chi_pad/C bond columns each.  Cost model, calibrated on the cfg3 branch launch (profiles/r05j_gc_sweep.txt):
        CTA-time per trajectory-step = C x sum over the passes of a tile (row blocks of <= 16 rows that share a PT
        block, per panel of 128 output columns) of (m-tiles x k-rows of this CTA + a fixed cost of the partial-product
        exchange, which doubles on 8-CTA clusters), times the waves of clusters."""
        if os.environ.get("ACEQD_SPLITK_GC") and not (tile_T or cluster):     # experiments: "G,C"
            tile_T, cluster = (int(x) for x in os.environ["ACEQD_SPLITK_GC"].split(","))
        _, blk_of_alpha = self.problem_handle(prob, pt)
        rows = np.bincount(blk_of_alpha)
        rows = rows[rows > 0]
        chi_pad = -(-pt.chi_max // 8) * 8
        n_pan = -(-chi_pad // 128)
        best = None
        for c in ((cluster,) if cluster else (2, 4, 8)):
            ovh = SPLITK_PASS_OVERHEAD * (2.0 if c == 8 else 1.0)
            for g in ((tile_T,) if tile_T else range(1, 17)):
                if g > max(1, n_traj) and not tile_T:
                    continue
                if not self.lib.aceqd_splitk_fit(prob.NL, chi_pad, g, c):
                    continue
                work = 0.0
                for r in rows:
                    mt = -(-(g * r) // 8)
                    work += n_pan * (mt * chi_pad / c + -(-mt // 2) * ovh)
                tiles = -(-n_traj // g)
                waves = -(-(tiles * c) // CLUSTER_CAPACITY[c])
                cost = waves * c * work / min(g, max(1, n_traj))
                if best is None or cost < best[2] - 1e-9:
                    best = (g, c, cost)
        if best is None:
            raise EngineError(f"NL={prob.NL}, chi={chi_pad}: no (tile, cluster) of the split-K kernel fits shared memory")
        return best

    def run_sweep(self, prob: Problem, pt: Optional[ProcessTensor], tables: np.ndarray,
                  grid: Tuple[float, float], t_start: float, n_steps: int, dt: float, *,
                  plan: Optional["_Plan"] = None, copy: bool = True, kernel: str = "auto",
                  t_eval: str = "half_mid", tile_T: Optional[int] = None) -> np.ndarray:
        """End-to-end sweep through the C ABI with HOST buffers: `tables[n_traj, 3, n_samples]`
        (x, y, rf drive samples of every trajectory; ideally :meth:`pinned_empty` memory) is
        copied to the device, the operators are built, the trajectories propagated and the
        outputs copied back.  Returns ``out[n_traj, n_steps+1, n_out]`` (a view of an
        engine-owned pinned buffer, valid until the next call, if ``copy=False``)."""
        if pt is None:
            pt = self._trivial(prob)
        hp, _ = self.problem_handle(prob, pt)
        hpt = self.pt_handle(pt)
        tables = np.ascontiguousarray(tables, dtype=np.complex128)
        n_traj, n_tab, n_samples = tables.shape
        if not 1 <= n_tab <= 3:
            raise ValueError("tables must be [n_traj, n_tab<=3 (x, y, rf), n_samples]")
        if plan is None:
            plan = self.plan_sweep(prob, pt, n_traj, n_steps, dt, t_start, n_traj, n_samples, grid,
                                   kernel=kernel, t_eval=t_eval, tile_T=tile_T)
        out = self._pinned_out(plan.batch.out_elems)
        plan.batch.tables = tables.ctypes.data
        plan.batch.n_tables = n_tab
        plan.batch.out = out.ctypes.data
        plan.batch.device_resident = 0
        _check(self.lib.aceqd_propagate_batch(self.ctx, hp, hpt, ctypes.byref(plan.batch)),
               "aceqd_propagate_batch")
        res = out.reshape(n_traj, n_steps + 1, prob.n_out)
        return res.copy() if copy else res

    def run_sweep_device(self, prob: Problem, pt: ProcessTensor, plan: "_Plan", tables_ptr: int,
                         out_ptr: int) -> None:
        """HBM-resident leg: drive tables and outputs are device pointers; asynchronous on the
        engine's stream."""
        hp, _ = self.problem_handle(prob, pt)
        hpt = self.pt_handle(pt)
        plan.batch.tables = tables_ptr
        plan.batch.out = out_ptr
        plan.batch.device_resident = 1
        _check(self.lib.aceqd_propagate_batch(self.ctx, hp, hpt, ctypes.byref(plan.batch)),
               "aceqd_propagate_batch(device)")

    def sync(self):
        _check(self.lib.aceqd_ctx_sync(self.ctx), "aceqd_ctx_sync")

    # -------------------------------------------------------------- planning
    def plan(self, prob: Problem, pt: ProcessTensor, jobs: Sequence[Job], *, kernel: str = "auto",
             t_eval: str = "half_mid", fork: bool = True, tile_T: Optional[int] = None,
             cluster: Optional[int] = None, trunk_kernel: Optional[str] = None):
        """Levels of trajectory descriptors for ``jobs`` (:mod:`pyaceqd_b200.planner`): ``(common, plan)``."""
        return self.plan_arrays(prob, pt, planner.arrays_from_jobs(prob, jobs), kernel=kernel, t_eval=t_eval, fork=fork,
                                tile_T=tile_T, cluster=cluster, trunk_kernel=trunk_kernel)

    def plan_arrays(self, prob: Problem, pt: ProcessTensor, arr: "planner.JobArrays", *, kernel: str = "auto",
                    t_eval: str = "half_mid", fork: bool = True, tile_T: Optional[int] = None,
                    cluster: Optional[int] = None, trunk_kernel: Optional[str] = None):
        chi_pad = -(-pt.chi_max // 8) * 8
        common = dict(prob=prob, pt=pt, dt=arr.dt, t0=arr.t0, off=T_EVAL[t_eval], packed=arr.packed, grid=arr.grid,
                      mats=arr.mats, chi_pad=chi_pad, kernel=resolve_kernel(kernel, prob.NL), tile_T=tile_T,
                      cluster=cluster, trunk_kernel=trunk_kernel, rho0s=arr.rho0s)
        return common, planner.plan_levels(arr, prob.n_out, fork=fork)

    def _materialise(self, common, lv: "planner.Level", out_off_of_traj, out_elems, n_slots=0, out_buf=None) -> _Plan:
        prob, pt = common["prob"], common["pt"]
        NL = prob.NL
        seqs = np.zeros(len(lv.seqs), dtype=SEQ_DT)
        seqs["set"], seqs["step0"], seqs["len"], seqs["first_has_prev"] = lv.seqs.T
        seq_base = np.zeros(len(seqs) + 1, dtype=np.int64)
        seq_base[1:] = np.cumsum(seqs["len"].astype(np.int64))
        entries = np.zeros(len(lv.entries), dtype=ENTRY_DT)
        if len(entries):
            for k, name in enumerate(("set", "step", "sb", "sa", "has_prev", "clamp")):
                entries[name] = lv.entries[:, k]
        n = lv.n_traj
        trajs = np.zeros(n, dtype=TRAJ_DT)
        trajs["ent0"] = seq_base[lv.seq] + lv.off
        trajs["out_off"] = out_off_of_traj
        for name in ("step0", "n_steps", "init_kind", "init_index", "n_ovr", "ovr_step", "ovr_ent", "out_from",
                     "snap_off", "snap_cnt", "snap_slot0"):
            trajs[name] = getattr(lv, name)
        # A uniform PT (one periodic slice, no initial block) is invariant under time translation: the step kernel
        # needs a trajectory's absolute start only to pick PT slices and closures, the per-row operators carry the
        # absolute times themselves (sequence / entry steps).  All trajectories are therefore rebased to start
        # together, so the members of a tile are active in the same rows (a tile of a (t, tau) map takes n_tau steps
        # instead of n_tau + T - 1; triangular sweeps tile by length alone).  Start row 1, not 0: a snapshot-started
        # trajectory reads the closure of the slice BEFORE its first row.
        if pt.n_initial == 0 and pt.n_repeat == 1 and os.environ.get("ACEQD_REBASE", "1") != "0":
            trajs["step0"] = 1
        # tiling: sort by (step0, n_steps) so that tiles are homogeneous in absolute time
        order = np.lexsort((trajs["n_steps"], trajs["step0"]))
        t_max = self.max_tile(NL, common["chi_pad"])
        kernel = common["kernel"]
        if len(lv.snap_steps) and kernel in ("small", "splitk"):
            # trunks (few trajectories that write snapshots): not a small-bond kernel job, and a single trajectory is
            # faster on the tile kernel's pass-split cluster than on the split-K cluster
            kernel = common.get("trunk_kernel") or "tile"
        if t_max < 1 and kernel in ("dmma", "tile"):
            kernel = "splitk"     # one trajectory does not fit a CTA: spread its bond columns over a cluster
        if t_max < 1 and kernel != "splitk":
            raise EngineError(f"NL={NL}, chi={common['chi_pad']} does not fit the step kernel's shared memory")
        T, C = self._tile_and_cluster(prob, pt, n, t_max, common["tile_T"], common.get("cluster"), kernel)
        n_tiles = -(-n // T)
        tile_traj = np.full(n_tiles * T, -1, dtype=np.int32)
        tile_traj[:n] = order
        mats = _c128(np.asarray(common["mats"]).reshape(-1, NL, NL)) if common["mats"] else np.zeros((0, NL, NL), complex)
        rho0 = _c128(common["rho0s"]).reshape(-1, NL)
        snap_steps = np.ascontiguousarray(lv.snap_steps, dtype=np.int32)
        out = out_buf if out_buf is not None else np.zeros(out_elems, dtype=np.complex128)
        packed = common["packed"]
        b = _Batch()
        b.dt, b.t0 = common["dt"], common["t0"]
        b.eval_off1, b.eval_off2 = common["off"]
        b.n_sets, b.n_tables, b.n_samples = packed.shape
        b.tab_t0, b.tab_dt = common["grid"]
        b.tables = packed.ctypes.data
        b.n_seq, b.seqs = len(seqs), seqs.ctypes.data
        b.n_entries, b.entries = len(entries), (entries.ctypes.data if len(entries) else None)
        b.n_mto_mats, b.mto_mats = len(mats), (mats.ctypes.data if len(mats) else None)
        b.n_rho0, b.rho0s = rho0.shape[0], rho0.ctypes.data
        b.n_traj, b.trajs = n, trajs.ctypes.data
        b.tile_T, b.n_tiles, b.tile_traj = T, n_tiles, tile_traj.ctypes.data
        b.n_snap_steps = len(snap_steps)
        b.snap_steps = snap_steps.ctypes.data if len(snap_steps) else None
        b.n_snap_slots = n_slots      # the pool of the WHOLE plan in every launch: later levels read earlier slots
        b.out_elems, b.out = out_elems, out.ctypes.data
        b.device_resident = 0
        b.kernel = KERNELS[kernel]
        b.cluster = C
        return _Plan(batch=b, keep=[seqs, entries, trajs, tile_traj, mats, rho0, snap_steps, packed, out],
                     out=out, out_off=np.asarray(out_off_of_traj), n_rows=trajs["n_steps"] + 1)

    # -------------------------------------------------------------- running
    def run_arrays(self, prob: Problem, pt: Optional[ProcessTensor], arr: "planner.JobArrays", *, kernel: str = "auto",
                   t_eval: str = "half_mid", fork: bool = True, tile_T: Optional[int] = None,
                   cluster: Optional[int] = None, trunk_kernel: Optional[str] = None, tail_reduce=None):
        """Propagate the jobs of ``arr``.  Returns ``(out, out_off, n_rows)``: job ``i`` owns ``out[out_off[i]:
        out_off[i] + n_rows[i] * n_out]``, rows x outputs -- or, with ``tail_reduce``, ``[n_jobs, n_pairs]``."""
        if pt is None:
            pt = self._trivial(prob)
        hp, _ = self.problem_handle(prob, pt)
        hpt = self.pt_handle(pt)
        common, pl = self.plan_arrays(prob, pt, arr, kernel=kernel, t_eval=t_eval, fork=fork, tile_T=tile_T,
                                      cluster=cluster, trunk_kernel=trunk_kernel)
        n_out = prob.n_out
        root_out = None
        for d, lv in enumerate(pl.levels[:-1]):
            rows = lv.n_steps + 1 - lv.out_from
            off = np.zeros(lv.n_traj, dtype=np.int64)
            off[1:] = np.cumsum(rows[:-1] * n_out)
            tp = self._materialise(common, lv, off, int(np.sum(rows * n_out)), n_slots=pl.n_slots)
            _check(self.lib.aceqd_propagate_batch(self.ctx, hp, hpt, ctypes.byref(tp.batch)),
                   "aceqd_propagate_batch(trunk)" if lv.group is not None else "aceqd_propagate_batch(fork level)")
            self._log_launch("trunk" if lv.group is not None else "fork", prob, common["chi_pad"], tp)
            if lv.group is not None:
                root_out = (tp.out, off)
        main = pl.levels[-1]
        # a branch writes its rows at the tail of its job's block
        traj_out_off = pl.out_off[main.job] + main.row0 * n_out
        if tail_reduce is not None and not len(pl.copies):
            pairs, spacing = tail_reduce
            ch = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
            red = np.zeros((main.n_traj, len(ch)), dtype=np.complex128)
            mp = self._materialise(common, main, traj_out_off, pl.out_elems, n_slots=pl.n_slots, out_buf=red.reshape(-1))
            mp.batch.n_reduce, mp.batch.reduce_ch = len(ch), ch.ctypes.data
            mp.batch.reduce_spacing, mp.batch.reduce_out = float(spacing), red.ctypes.data
            mp.batch.out_elems = pl.out_elems
            _check(self.lib.aceqd_propagate_batch(self.ctx, hp, hpt, ctypes.byref(mp.batch)),
                   "aceqd_propagate_batch(tail_reduce)")
            self._log_launch("main", prob, common["chi_pad"], mp)
            return red
        mp = self._materialise(common, main, traj_out_off, pl.out_elems, n_slots=pl.n_slots)
        _check(self.lib.aceqd_propagate_batch(self.ctx, hp, hpt, ctypes.byref(mp.batch)),
               "aceqd_propagate_batch")
        self._log_launch("main", prob, common["chi_pad"], mp)
        out = mp.out
        for (job, n_copy, ti, row_first) in pl.copies:
            a = root_out[1][ti] + row_first * n_out
            out[pl.out_off[job]: pl.out_off[job] + n_copy * n_out] = root_out[0][a: a + n_copy * n_out]
        if tail_reduce is not None:      # rows partly on the trunk: reduce on the host (same arithmetic)
            return np.asarray([tail_trapezoid(out[o: o + r * n_out].reshape(r, n_out).T, *tail_reduce)
                               for o, r in zip(pl.out_off, pl.n_rows)])
        return out, pl.out_off, pl.n_rows

    def run_jobs(self, prob: Problem, pt: Optional[ProcessTensor], jobs: Sequence[Job], *,
                 kernel: str = "auto", t_eval: str = "half_mid", fork: bool = True,
                 tile_T: Optional[int] = None, cluster: Optional[int] = None,
                 trunk_kernel: Optional[str] = None, tail_reduce=None) -> List[np.ndarray]:
        """Propagate `jobs`; returns one ``[n_out, n_steps+1]`` complex array per job (its last ``tail_rows`` columns
        if the job asks for a tail).

        ``tail_reduce = (pairs, spacing)`` fuses the consumer's tau integral into the launch (reference
        ``pol_entanglement/G2.py:507-533``): ``pairs = [(ch_tau, ch_zero), ...]`` output channels; the outputs stay in
        HBM and per job one ``[len(pairs)]`` array comes back -- ``spacing`` times the trapezoid over the job's kept rows,
        the first row (tau = 0) read from ``ch_zero``, the others from ``ch_tau``."""
        res = self.run_arrays(prob, pt, planner.arrays_from_jobs(prob, jobs), kernel=kernel, t_eval=t_eval, fork=fork,
                              tile_T=tile_T, cluster=cluster, trunk_kernel=trunk_kernel, tail_reduce=tail_reduce)
        if tail_reduce is not None:
            return [r.copy() for r in res]
        out, out_off, n_rows = res
        n_out = prob.n_out
        return [np.ascontiguousarray(out[o: o + r * n_out].reshape(r, n_out).T) for o, r in zip(out_off, n_rows)]

    def _log_launch(self, kind: str, prob: Problem, chi_pad: int, plan: "_Plan"):
        if not self.record_timings:
            return
        step_ms, op_ms = self.last_timings()
        names = self.last_kernels()
        self.timing_log.append(dict(kind=kind, step_ms=step_ms, opbuild_ms=op_ms, NL=prob.NL, chi_pad=chi_pad,
                                    step_kernel=names["step"], opbuild_kernel=names["opbuild"],
                                    kernel=int(plan.batch.kernel),
                                    n_traj=int(plan.batch.n_traj), tile_T=int(plan.batch.tile_T),
                                    cluster=int(plan.batch.cluster),
                                    n_tiles=int(plan.batch.n_tiles),
                                    traj_steps=int(np.sum(np.asarray(plan.n_rows) - 1))))

    def _trivial(self, prob: Problem) -> ProcessTensor:
        key = ("trivial", len(prob.cls_keys))
        if not hasattr(self, "_triv"):
            self._triv = {}
        if key not in self._triv:
            self._triv[key] = trivial_pt(n_cls=len(prob.cls_keys))
        return self._triv[key]


_default_engines: Dict[int, Engine] = {}


def default_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]
