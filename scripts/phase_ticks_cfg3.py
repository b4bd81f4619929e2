#!/usr/bin/env python
"""Per-phase cycle breakdown (CTA 0) of the cfg3 G2 map's BRANCH launch (2-CTA clusters) or TRUNK launch (8-CTA cluster).
usage: ACEQD_TICK_CLUSTER=2 scripts/phase_ticks_cfg3.py"""
import ctypes, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.two_time.correlations import three_op_two_time

n_t, dt = 256, 0.25
eng = default_engine(0)
lib = eng.lib
lib.aceqd_debug_phase_ticks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
pt = synthetic_pt(128, 9, dt=dt, seed=1234, kind="unitary", scale=0.999)
f = os.path.join(tempfile.mkdtemp(), "pt.pt"); pt.save(f)
pulse = ChirpedPulse(tau_0=5.0, e_start=-2.0, alpha=0, t0=20.0, e0=5.0, polar_x=1.0)
t_axis = np.round(dt * np.arange(n_t), 6)
opts = {"lindblad": True, "phonons": True, "pt_file": f, "delta_b": 4.0}
run = lambda: three_op_two_time(biexciton, t_axis, pulse, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4",
                                tau_max=n_t * dt, dt=dt, options=dict(opts))
run()
lib.aceqd_debug_phase_ticks(eng.ctx, 1, None)
run()
t = np.zeros(8, dtype=np.int64)
lib.aceqd_debug_phase_ticks(eng.ctx, 0, t.ctypes.data)
names = ["wait W/OV (+ row exchange wait)", "outputs", "system product + barrier", "GEMM main loops", "barrier after passes",
         "closure sums + barrier", "step tail (push, peers)", "after each pass (barrier, push, epilogue)"]
tot = t.sum()
for n, v in zip(names, t):
    print(f"{n:44s} {v:14d} cycles  {100.0 * v / max(tot, 1):5.1f}%")
print("total", tot, "cycles; cluster filter", os.environ.get("ACEQD_TICK_CLUSTER"))
