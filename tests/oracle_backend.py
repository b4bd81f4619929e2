"""Test double: an engine whose ``run_jobs`` is the CPU oracle.

Lets the ``-m "not gpu"`` suite exercise every batch issuer (two_time, pol_entanglement, timebin,
rabi / tpe sweeps) end to end through the deferred-execution path, and gives the GPU suite a
workflow-level parity reference.  Test infrastructure only (it imports ``oracle/``)."""
import contextlib
import copy

import numpy as np
import scipy.linalg

import oracle
from pyaceqd_b200.process_tensor import trivial_pt


class OracleEngine:
    def __init__(self):
        self.calls = []          # (n_jobs, total steps) per batch: asserts on batching behaviour

    def run_jobs(self, prob, pt, jobs, **kw):
        if pt is None:
            pt = trivial_pt(n_cls=len(prob.cls_keys))
        self.calls.append((len(jobs), sum(j.n_steps for j in jobs)))
        out = []
        for j in jobs:
            pj = prob
            if getattr(j, "rho0", None) is not None:      # per-job initial state (dynamical maps)
                pj = copy.copy(prob)
                pj.rho0 = np.asarray(j.rho0, dtype=complex).reshape(-1)
            full = oracle.propagate(pj, pt, j, t_eval=kw.get("t_eval", "half_mid"))
            out.append(np.ascontiguousarray(full[:, -j.tail_rows:]) if j.tail_rows else full)
        if kw.get("tail_reduce") is not None:      # host statement of the fused tail reduction
            from pyaceqd_b200.engine import tail_trapezoid
            return [tail_trapezoid(o, *kw["tail_reduce"]) for o in out]
        return out

    def run_arrays(self, prob, pt, arr, **kw):
        """The array route (``Engine.run_arrays``): every row of the job table becomes an UN-forked oracle run, its
        merged operator products re-issued as multi-time operators at their steps."""
        from pyaceqd_b200.jobs import FieldTable, Job
        from pyaceqd_b200.problem import MTO
        jobs = []
        tabs_of_set = {}
        for j in range(arr.n_jobs):
            s = int(arr.set_id[j])
            if s not in tabs_of_set:
                tabs_of_set[s] = {pol: FieldTable(arr.grid[0], arr.grid[1], arr.packed[s, k].copy())
                                  for k, pol in enumerate(("x", "y", "rf")) if np.any(arr.packed[s, k] != 0)}
            t_start = arr.t0 + int(arr.shift[j]) * arr.dt
            mtos = []
            for e in range(int(arr.n_ev[j])):
                t = t_start + int(arr.ev_step[j, e]) * arr.dt
                if arr.ev_sb[j, e] >= 0:
                    mtos.append(MTO(arr.mats[arr.ev_sb[j, e]], t, True))
                if arr.ev_sa[j, e] >= 0:
                    mtos.append(MTO(arr.mats[arr.ev_sa[j, e]], t, False))
            jb = Job(t_start, t_start + int(arr.n_steps[j]) * arr.dt, arr.dt, tables=tabs_of_set[s], mtos=mtos,
                     tail_rows=int(arr.tail[j]))
            jb.table_len = int(arr.clamp[j])
            if arr.r0[j]:
                jb.rho0 = arr.rho0s[arr.r0[j]]
            jobs.append(jb)
        outs = self.run_jobs(prob, pt, jobs, **kw)
        if kw.get("tail_reduce") is not None:
            return np.asarray(outs)
        n_rows = np.asarray([o.shape[1] for o in outs], dtype=np.int64)
        out_off = np.zeros(len(outs), dtype=np.int64)
        out_off[1:] = np.cumsum(n_rows[:-1] * prob.n_out)
        return np.concatenate([o.T.reshape(-1) for o in outs]), out_off, n_rows

    def expm(self, mats):
        a = np.asarray(mats, dtype=complex)
        a = a[None] if a.ndim == 2 else a
        return np.array([scipy.linalg.expm(m) for m in a])


@contextlib.contextmanager
def oracle_backend():
    """Route ``default_engine()`` (looked up at call time by ``general_system.run_requests``) to the
    oracle for the duration of the block."""
    import pyaceqd_b200.engine as eng_mod
    saved = eng_mod.default_engine
    eng = OracleEngine()
    eng_mod.default_engine = lambda device=None: eng
    try:
        yield eng
    finally:
        eng_mod.default_engine = saved


def run_programs_numpy(mats, v0, seg_off, segs, w=None, n_emit_max=0, want_final=False):
    """Plain interpreter of chain programs (same contract as ``Engine.tlmap_run``)."""
    mats = np.asarray(mats, dtype=complex)
    NL = mats.shape[-1]
    n_chains = len(v0)
    w = np.asarray(w, dtype=complex).reshape(-1, NL) if w is not None else np.zeros((0, NL), complex)
    out = np.zeros((n_chains, n_emit_max, len(w)), dtype=complex) if n_emit_max and len(w) else None
    fin = np.zeros((n_chains, NL), dtype=complex) if want_final else None
    for c in range(n_chains):
        v = np.array(v0[c], dtype=complex)
        e = 0
        for s in segs[seg_off[c]:seg_off[c + 1]]:
            for k in range(int(s["count"])):
                v = mats[int(s["start"]) + k * int(s["stride"])] @ v
                if s["emit"] and out is not None and e < n_emit_max:
                    out[c, e] = w @ v
                    e += 1
        if fin is not None:
            fin[c] = v
    return out, fin


OracleEngine.tlmap_run = staticmethod(run_programs_numpy)
