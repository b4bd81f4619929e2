"""Where does the host time of a cfg3 G2 map go?  (run on a GPU box)"""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyaceqd_b200 import engine as E
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.general_system import general_system as gs
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.two_time.correlations import three_op_two_time
import pyaceqd_b200.batch as B

n_t, dt = 256, 0.25
eng = default_engine(0)
pt = synthetic_pt(128, 9, dt=dt, seed=1234, kind="unitary", scale=0.999)
f = os.path.join(tempfile.mkdtemp(), "pt.pt"); pt.save(f)
pulse = ChirpedPulse(tau_0=5.0, e_start=-2.0, alpha=0, t0=20.0, e0=5.0, polar_x=1.0)
t_axis = np.round(dt * np.arange(n_t), 6)
opts = {"lindblad": True, "phonons": True, "pt_file": f, "delta_b": 4.0}
T = {}
def wrap(obj, name, key):
    fn = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); T[key] = T.get(key, 0.0) + time.perf_counter() - t; return r
    setattr(obj, name, w)
from pyaceqd_b200 import planner as P
wrap(E.Engine, "plan_arrays", "plan"); wrap(E.Engine, "_materialise", "materialise"); wrap(E.Engine, "run_arrays", "run_arrays")
wrap(P, "arrays_from_sweep", "arrays_from_sweep"); wrap(gs, "run_sweep_arrays", "run_sweep_arrays")
wrap(gs, "run_requests", "run_requests"); wrap(B.BatchExecutor, "submit", "submit")
orig = eng.lib.aceqd_propagate_batch
class L:
    def __getattr__(self, n): return getattr(orig_lib, n)
orig_lib = eng.lib
def prop(*a):
    t = time.perf_counter(); r = orig(*a); T["propagate_batch"] = T.get("propagate_batch", 0.0) + time.perf_counter() - t; return r
run = lambda: three_op_two_time(biexciton, t_axis, pulse, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4", tau_max=n_t * dt, dt=dt, options=dict(opts))
run(); run()
eng.lib.aceqd_propagate_batch = prop
for rep in range(3):
    T.clear()
    t = time.perf_counter(); run(); w = time.perf_counter() - t
    print("wall %.2f ms: " % (1e3 * w) + ", ".join("%s %.2f" % (k, 1e3 * v) for k, v in sorted(T.items())))

# a 256 x 256 triangular sweep (four_time shape): planning time only
import numpy as np
n = 256
t1 = np.round(dt * (1 + np.arange(n)), 6)
ii, jj = np.triu_indices(n)
a, b = t1[ii], t1[jj]
tb = 75.0
prob = gs._problem_cache[next(k for k in gs._problem_cache if not (isinstance(k, tuple) and k and k[0] == "dynmap"))]
parsed = prob.parse_mtos([{"operator": "|3><1|_4", "applyFrom": "_right", "time": 0.0}, {"operator": "|1><3|_4", "applyFrom": "_left", "time": 0.0},
                          {"operator": "|0><1|_4", "applyFrom": "_left", "time": 0.0}])
from pyaceqd_b200.jobs import FieldTable
tabs = {"x": FieldTable(0.0, dt, pulse.get_total(dt * np.arange(int((2 * n * dt + tb) / dt) + 2)))}
for rep in range(3):
    t = time.perf_counter()
    arr = P.arrays_from_sweep(prob, dt=dt, t_start=0.0, t_end=b + tb, superops=[m.superop for m in parsed], before=[m.before for m in parsed],
                              mto_times=np.stack([a, b, a + tb], axis=1), tails=1, tables=tabs)
    t2 = time.perf_counter()
    pl = P.plan_levels(arr, prob.n_out)
    t3 = time.perf_counter()
    print("triangular 256x256 (%d jobs): arrays %.2f ms, plan_levels %.2f ms, levels %s, snapshot slots %d" %
          (arr.n_jobs, 1e3 * (t2 - t), 1e3 * (t3 - t2), [l.n_traj for l in pl.levels], pl.n_slots))
