#!/usr/bin/env python
"""Per-phase cycle breakdown (CTA 0, thread 0) of the split-K cluster kernel on the cfg3 branch launch.
usage: ACEQD_SPLITK_GC=8,4 scripts/phase_ticks_splitk.py"""
import ctypes, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ACEQD_KERNEL"] = "splitk"
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.two_time.correlations import three_op_two_time

n_t, dt = 256, 0.25
eng = default_engine(0)
eng.record_timings = True
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
os.environ["ACEQD_TICK_CLUSTER"] = os.environ.get("ACEQD_SPLITK_GC", "16,8").split(",")[1]
lib.aceqd_debug_phase_ticks(eng.ctx, 1, None)
eng.timing_log.clear()
run()
t = np.zeros(8, dtype=np.int64)
lib.aceqd_debug_phase_ticks(eng.ctx, 0, t.ctypes.data)
names = ["phase A (closure exchange wait, outputs)", "system product", "GEMM main loops (+ chunk waits)", "wait: slot consumed (rempty)",
         "partial-product stores (st.async) + barrier", "to the reduce of the previous pass", "reduce: wait partials (rfull)",
         "reduce: sums, barrier, signal, loop"]
tot = t.sum()
main = [l for l in eng.timing_log if l["kind"] == "main"][-1]
print(main["step_kernel"], "step_ms", main["step_ms"])
for n, v in zip(names, t):
    print(f"{n:48s} {v:14d} cycles  {100.0 * v / max(tot, 1):5.1f}%")
print("total", tot, "cycles")
