"""Step-kernel throughput on the OTHER shapes BASELINE.json names (aligned pulse-area sweeps, tile kernel):
biexciton NL=16 chi=128 (cfg3-like), six-level NL=36 chi=128 (cfg4), five-level NL=25 chi=256 (cfg5).
One JSON line per shape: device times from the library's own events.  Run on a GPU box."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import biexciton_problem, make_tables, sixls_problem
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.jobs import Job
from pyaceqd_b200.problem import build_problem
from pyaceqd_b200.process_tensor import synthetic_pt
from pyaceqd_b200.pulses import ChirpedPulse


def fivels_problem():
    lb = [["|0><1|_5", 0.01], ["|0><2|_5", 0.01], ["|1><4|_5", 0.012], ["|2><4|_5", 0.012]]
    return build_problem(system_op=["-4.0*|4><4|_5", "-0.05*|1><1|_5", "0.05*|2><2|_5", "-0.3*|3><3|_5",
                                    "0.02*(|1><3|_5 + |3><1|_5)"],
                         boson_op="1*(|1><1|_5 + |2><2|_5 + |3><3|_5) + 2*|4><4|_5", initial="|0><0|_5", lindblad_ops=lb,
                         interaction_ops=[["|1><0|_5+|4><1|_5", "x"], ["|2><0|_5+|4><2|_5", "y"]],
                         output_ops=["|0><0|_5", "|1><1|_5", "|4><4|_5"])


eng = default_engine(0)
eng.record_timings = True
peak = eng.fp64_peak("dmma", 20000)
shapes = [("biexciton", biexciton_problem(outputs=["|1><1|_4", "|3><3|_4"]), 128, 2048, 100, 0.25),
          ("sixls", sixls_problem(), 128, 1184, 60, 0.1),
          ("fivels", fivels_problem(), 256, 592, 60, 0.1)]
kernel = "dmma"
tile = None
args = [a for a in sys.argv[1:]]
for a in list(args):
    if a.startswith("--kernel="):
        kernel = a.split("=", 1)[1]
        args.remove(a)
    if a.startswith("--tile="):
        tile = int(a.split("=", 1)[1])
        args.remove(a)
ticks = "--ticks" in args      # per-phase cycle clock of CTA 0 (aceqd_debug_phase_ticks)
if ticks:
    args.remove("--ticks")
    import ctypes
    eng.lib.aceqd_debug_phase_ticks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
if args:
    shapes = [s for s in shapes if s[0] in args]
for name, prob, chi, n_traj, n_steps, dt in shapes:
    pt = synthetic_pt(chi, len(prob.cls_keys), kind="unitary", scale=0.999)
    jobs = []
    for a in np.linspace(0.5, 12.0, n_traj):
        p = ChirpedPulse(tau_0=3.0, e_start=-2.0, alpha=0, t0=4.0, e0=a, polar_x=0.8)
        jobs.append(Job(0.0, n_steps * dt, dt, tables=make_tables([p], 0.0, n_steps * dt, dt), tail_rows=1))
    eng.run_jobs(prob, pt, jobs, kernel=kernel, tile_T=tile)
    eng.timing_log.clear()
    if ticks:
        eng.lib.aceqd_debug_phase_ticks(eng.ctx, 1, None)
    t = time.perf_counter(); out = eng.run_jobs(prob, pt, jobs, kernel=kernel, tile_T=tile); wall = time.perf_counter() - t
    if ticks:
        tk = np.zeros(8, dtype=np.int64)
        eng.lib.aceqd_debug_phase_ticks(eng.ctx, 0, tk.ctypes.data)
        names = ["wait W/OV", "outputs", "system product + barrier", "GEMM main loops", "barrier after passes",
                 "closure sums + barrier", "step tail", "after each pass (barrier, push, epilogue)"]
        for nm, v in zip(names, tk):
            print("  %-44s %12d cycles %5.1f%%" % (nm, v, 100.0 * v / max(1, tk.sum())))
    l = eng.timing_log[-1]
    NL = prob.NL
    fl = 8.0 * NL * chi * (2 * NL + chi) * n_traj * n_steps
    print(json.dumps(dict(shape=name, kernel=l['step_kernel'], opbuild_kernel=l['opbuild_kernel'], NL=NL, chi=chi, n_cls=len(prob.cls_keys), n_traj=n_traj, n_steps=n_steps,
                          tile_T=l["tile_T"], cluster=l["cluster"], step_ms=l["step_ms"], opbuild_ms=l["opbuild_ms"],
                          wall_ms=1e3 * wall, tflops=fl / (l["step_ms"] * 1e-3) / 1e12,
                          frac_of_dmma_peak=fl / (l["step_ms"] * 1e-3) / 1e12 / peak, dmma_peak=peak,
                          traj_steps_per_s=n_traj * n_steps / (l["step_ms"] * 1e-3),
                          traj_steps_per_s_with_operators=n_traj * n_steps / ((l["step_ms"] + l["opbuild_ms"]) * 1e-3))),
          flush=True)
