#!/usr/bin/env python
"""The cfg5-shaped triangular (t1, t2) sweep of bench.py's `strong` record at larger grids (one GPU).
usage: scripts/strong_sizes.py 48 96 192"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyaceqd_b200.engine import default_engine
eng = default_engine(0)
eng.record_timings = True
for n_t in [int(a) for a in sys.argv[1:]] or [48]:
    eng.timing_log.clear()
    t = time.perf_counter()
    rec = bench.measure_strong(eng, 1, 0, None, n_t=n_t)
    rec["n_t"] = n_t
    rec["script_wall_s_incl_warmup"] = time.perf_counter() - t
    half = eng.timing_log[len(eng.timing_log) // 2:]       # the timed run (second of two)
    rec["kernel_ms_sum"] = sum(l["step_ms"] + l["opbuild_ms"] for l in half)
    rec["launches"] = [(l["step_kernel"], round(l["step_ms"], 3), round(l["opbuild_ms"], 3)) for l in half]
    print(json.dumps(rec), flush=True)
