"""Diagnostic: where does the NL=36 / chi=128 MTO batch deviate from the oracle?"""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import oracle
from helpers import make_tables, sixls_problem, biexciton_problem
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse

eng = default_engine(0)
eng.record_timings = True
prob = sixls_problem()
dt = 0.1
p1 = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=1.5, e0=5.0, polar_x=0.7)
tabs = make_tables([p1], 0.0, 6.0, dt)


def jobs_of(mto, n=6):
    jobs = []
    for i in range(n):
        t1 = 0.3 * i
        mt = prob.parse_mtos([{"operator": "|0><1|_6", "applyFrom": "_left", "time": t1},
                              {"operator": "|2><0|_6", "applyFrom": "_right", "time": t1}]) if mto else []
        jobs.append(Job(0.0, t1 + 2.0, dt, tables=tabs, mtos=mt))
    return jobs


for chi, ns in ((40, 1), (64, 1), (128, 1), (128, 2)):
    pt = synthetic_pt(chi, len(prob.cls_keys), n_slices=ns, kind="unitary", scale=0.999)
    print("chi", chi, "slices", ns, "max_tile", eng.max_tile(prob.NL, -(-chi // 8) * 8), flush=True)
    for mto in (False, True):
        jobs = jobs_of(mto)
        refs = [oracle.propagate(prob, pt, j) for j in jobs]
        for kw in (dict(kernel="check", fork=False), dict(kernel="check", fork=True), dict(kernel="dmma", fork=False, cluster=1, tile_T=1),
                   dict(kernel="dmma", fork=False, cluster=1), dict(kernel="dmma", fork=False), dict(kernel="dmma", fork=True, cluster=1),
                   dict(kernel="dmma", fork=True), dict(kernel="dmma", fork=False, cluster=2, tile_T=2),
                   dict(kernel="dmma", fork=False, cluster=4, tile_T=1), dict(kernel="dmma", fork=False, cluster=8, tile_T=1)):
            eng.timing_log.clear()
            try:
                got = eng.run_jobs(prob, pt, jobs, **kw)
                dev = [float(np.abs(g - r).max()) for g, r in zip(got, refs)]
                first_bad = [int(np.argmax(np.abs(g - r).max(axis=0) > 1e-9)) for g, r in zip(got, refs)]
                print("  mto", mto, kw, "max dev %.2e" % max(dev), "per job", ["%.0e" % d for d in dev], "first bad row", first_bad,
                      [(l["kind"], l["step_kernel"]) for l in eng.timing_log], flush=True)
            except Exception as exc:   # noqa
                print("  mto", mto, kw, "ERROR", exc, flush=True)
