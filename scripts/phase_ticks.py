#!/usr/bin/env python
"""Per-phase cycle breakdown of the step kernel's CTA 0 (debug clock, aceqd_debug_phase_ticks) on a cfg2-style sweep.
usage: scripts/phase_ticks.py [chi] [n_area] [n_det] [n_steps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 128
na = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 400
from pyaceqd_b200.engine import default_engine
eng = default_engine(0)
lib = eng.lib
lib.aceqd_debug_phase_ticks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
dt = 0.1
prob, pt, tables = bench.make_workload(chi, na, nd, ns, dt)
tile = int(os.environ.get("TILE", "0")) or None
out = eng.run_sweep(prob, pt, tables, (0.0, dt), 0.0, ns, dt, copy=False, tile_T=tile)
lib.aceqd_debug_phase_ticks(eng.ctx, 1, None)
out = eng.run_sweep(prob, pt, tables, (0.0, dt), 0.0, ns, dt, copy=False, tile_T=tile)
t = np.zeros(8, dtype=np.int64)
lib.aceqd_debug_phase_ticks(eng.ctx, 0, t.ctypes.data)
names = ["wait W/OV", "outputs", "system product + barrier", "GEMM main loops", "barrier after passes",
         "closure sums + barrier", "step tail", "pass epilogues (arrive, wait, write)"]
tot = t.sum()
for n, v in zip(names, t):
    print(f"{n:40s} {v:14d} cycles  {100.0 * v / max(tot, 1):5.1f}%")
print("total", tot, "cycles; step kernel", eng.last_timings()[0], "ms")
