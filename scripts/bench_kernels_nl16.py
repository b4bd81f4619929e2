"""Persistent (shared-memory state, clusters) vs step-synchronous streaming kernel on an ALIGNED batch
of biexciton trajectories (NL = 16, chi = 128): the shape of a two-photon-excitation area sweep
(four_level_system/tpe_rotations.py:182-207).  Run on a GPU box."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import biexciton_problem, make_tables
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse

n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
eng = default_engine(0)
eng.record_timings = True
prob = biexciton_problem(outputs=["|1><1|_4", "|3><3|_4"])
pt = synthetic_pt(128, len(prob.cls_keys), kind="unitary", scale=0.999)
dt = 0.25
jobs = []
for a in np.linspace(0.5, 12.0, n_traj):
    p = ChirpedPulse(tau_0=3.0, e_start=-2.0, alpha=0, t0=10.0, e0=a)
    jobs.append(Job(0.0, n_steps * dt, dt, tables=make_tables([p], 0.0, n_steps * dt, dt), tail_rows=1))
res = {}
for kern in ("dmma", "stream"):
    eng.run_jobs(prob, pt, jobs, kernel=kern)
    eng.timing_log.clear()
    t = time.perf_counter(); out = eng.run_jobs(prob, pt, jobs, kernel=kern); wall = time.perf_counter() - t
    l = eng.timing_log[-1]
    fl = 8.0 * 16 * 128 * (32 + 128) * n_traj * n_steps
    res[kern] = dict(step_ms=l["step_ms"], opbuild_ms=l["opbuild_ms"], wall_ms=1e3 * wall, tile_T=l["tile_T"], cluster=l["cluster"],
                     tflops=fl / (l["step_ms"] * 1e-3) / 1e12, traj_steps_per_s=n_traj * n_steps / (l["step_ms"] * 1e-3),
                     final=float(np.real(out[7][0, -1])))
print(json.dumps({"n_traj": n_traj, "n_steps": n_steps, "NL": 16, "chi": 128, **res}))
