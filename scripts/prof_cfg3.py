"""cProfile of the cfg3 host path (run on a GPU box): where does the wall time of
three_op_two_time go besides the kernels?"""
import cProfile, pstats, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.two_time.correlations import three_op_two_time

n_t, dt = 256, 0.25
eng = default_engine(0)
pt = synthetic_pt(128, 9, dt=dt, seed=1234)
f = os.path.join(tempfile.mkdtemp(), "pt.pt"); pt.save(f)
pulse = ChirpedPulse(tau_0=5.0, e_start=-2.0, alpha=0, t0=20.0, e0=5.0, polar_x=1.0)
t_axis = np.round(dt * np.arange(n_t), 6)
opts = {"lindblad": True, "phonons": True, "pt_file": f, "delta_b": 4.0}
run = lambda: three_op_two_time(biexciton, t_axis, pulse, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4",
                                tau_max=n_t * dt, dt=dt, options=dict(opts))
run(); run()
t = time.perf_counter(); run(); print("wall ms", 1e3 * (time.perf_counter() - t))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
