"""Test double: an engine whose ``run_jobs`` is the CPU oracle.

Lets the ``-m "not gpu"`` suite exercise every batch issuer (two_time, pol_entanglement, timebin,
rabi / tpe sweeps) end to end through the deferred-execution path, and gives the GPU suite a
workflow-level parity reference.  Test infrastructure only (it imports ``oracle/``)."""
import contextlib

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
            full = oracle.propagate(prob, pt, j, t_eval=kw.get("t_eval", "half_mid"))
            out.append(np.ascontiguousarray(full[:, -j.tail_rows:]) if j.tail_rows else full)
        return out

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
