#!/usr/bin/env python
"""Benchmark of the PT-propagation hot path (BASELINE.json metric: trajectory-steps/sec at PT
bond dimension chi).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One bench "step" = one pass of the hot path over one batch: the cfg2 pulse-area x detuning
sweep of SURVEY 8d (two-level dot, synthetic PT of exact bond dimension chi, 64 x 64 = 4096
pulses, 400 time steps of dt = 0.1 ps).  A trajectory-step = half step exp(L dt/2) -> PT slice
-> half step -> closure/outputs for one trajectory.  Per rank the work is fixed (weak scaling):
rank r sweeps its own detuning window; the only collective is the final all-gather of results.

`value`  : whole-job trajectory-steps/s with drive tables and outputs resident in HBM
           (operator builder + step kernel, CUDA events, max over ranks).
`e2e`    : same metric through Engine.run_sweep -> aceqd_propagate_batch with HOST (pinned)
           buffers: H2D of the drive tables and D2H of all outputs inside the timed region.
`roofline`: step kernel only, algorithmic flops 8*NL*chi*(2*NL+chi) per trajectory-step
            over its CUDA-event duration, against the FP64 DMMA peak measured by the
           library's own register-resident micro-benchmark in this run (MEASURED_PEAKS.json has
           no FP64 figure).
`--impl reference`: the CPU restatement of the reference path (oracle/oracle_c.c, "port": the
           reference's own implementation is the external ACE binary, absent here) on all host
           threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBAR = 0.6582119569
METRIC = "trajectory-steps/sec at PT bond dim chi"
UNIT = "trajectory-steps/s"


def make_workload(chi, n_area, n_det, n_steps, dt, rank=0):
    """cfg2 of SURVEY 8d: areas linspace(0,30,n_area) (units of pi), detunings
    linspace(-5,5,n_det) meV (shifted by 10 meV per rank), Gaussian tau=5 ps at t0=20 ps."""
    from pyaceqd_b200.problem import build_problem
    from pyaceqd_b200.process_tensor import synthetic_pt
    prob = build_problem(boson_op="1.000*|1><1|_2", initial="|0><0|_2", lindblad_ops=[["|0><1|_2", 0.01]],
                         interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2"])
    pt = synthetic_pt(chi, len(prob.cls_keys), dt=dt, seed=1234, kind="unitary", scale=0.999)   # keeps the signal O(1) over all steps
    t = dt * np.arange(n_steps)                       # pulse-file grid np.arange(t_start, t_end, dt)
    areas = np.linspace(0.0, 30.0, n_area)
    dets = np.linspace(-5.0, 5.0, n_det) + 10.0 * rank
    tau, t0 = 5.0, 20.0
    env = np.exp(-0.5 * ((t - t0) / tau) ** 2) / (np.sqrt(2 * np.pi) * tau)
    f = (areas[:, None, None] * env[None, None, :]) * np.exp(-1j * (dets[None, :, None] / HBAR) * (t - t0)[None, None, :])
    f = f.reshape(n_area * n_det, 1, n_steps)
    f = np.round(f.real, 8) + 1j * np.round(f.imag, 8)   # %.8f pulse-file quantisation
    return prob, pt, np.ascontiguousarray(f)


def flops_per_step(NL, chi):
    return 8.0 * NL * chi * (2 * NL + chi)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.t_begin is not None and not (self.t_begin <= ts <= (self.t_end or ts) + 0.15):
                continue
            c = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(c[0]))
                mx = max(mx, float(c[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        # the median over samples under load (above the idle clock)
        load = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_rate(prob, pt, tables, n_steps, dt, seconds=10.0, threads=0):
    """Time the C oracle on a bounded sample of the same workload; returns (rate, sample str, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    from pyaceqd_b200.jobs import FieldTable, Job
    if threads <= 0:
        # all host threads: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently
        # turn the CPU baseline into a single-thread run
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = threads

    def run(n):
        idx = np.linspace(0, tables.shape[0] - 1, n).astype(int)
        jobs = [Job(0.0, n_steps * dt, dt, tables={"x": FieldTable(0.0, dt, tables[i, 0])}) for i in idx]
        t = time.perf_counter()
        oracle_c.propagate_sweep(prob, pt, jobs, n_threads=threads)
        return time.perf_counter() - t

    probe = max(cores, 8)
    t_probe = run(probe)
    n = int(min(tables.shape[0], max(probe, probe * seconds / max(t_probe, 1e-3))))
    n = max(cores, n // cores * cores)
    el = run(n)
    return n * n_steps / el, f"{n} of {tables.shape[0]} trajectories x {n_steps} steps ({el:.1f} s)", cores


def measure_g2_grid(local, n_t=256, chi=128, steps=3, warmup=2, dist=None, cpu=True, peak_dmma=None):
    """cfg3 of SURVEY 8d -- the north-star shape: biexciton (NL=16, 9 coupling classes) + synthetic PT chi=128,
    two-photon excitation pulse, G2(t,tau) on an n_t x n_t grid through the public workflow ``three_op_two_time``
    (reference two_time/correlations.py:227-270): one trunk + n_t forked branches of n_t steps.  Returns a dict: grid
    wall time, branch / trunk kernel times (CUDA events of the library), roofline of the branch launch, and the CPU
    restatement extrapolated to the reference's un-forked trajectory-steps."""
    import tempfile
    from pyaceqd_b200.engine import default_engine
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.problem import build_problem
    from pyaceqd_b200.process_tensor import synthetic_pt
    from pyaceqd_b200.pulses import ChirpedPulse
    from pyaceqd_b200.two_time.correlations import three_op_two_time

    dt, tau_max = 0.25, 0.25 * n_t
    eng = default_engine(local)
    eng.record_timings = True
    if peak_dmma is None:
        peak_dmma = eng.fp64_peak("dmma", 20000)
    pt = synthetic_pt(chi, 9, dt=dt, seed=1234, kind="unitary", scale=0.999)
    tmp = tempfile.mkdtemp(prefix="aceqd_bench_")
    pt_file = os.path.join(tmp, "synthetic_chi%d.pt" % chi)
    pt.save(pt_file)
    pulse = ChirpedPulse(tau_0=5.0, e_start=-2.0, alpha=0, t0=20.0, e0=5.0, polar_x=1.0)
    t_axis = np.round(dt * np.arange(n_t), 6)
    opts = {"lindblad": True, "phonons": True, "pt_file": pt_file, "delta_b": 4.0, "gamma_e": 0.01, "gamma_b": 0.01}

    def one_pass():
        eng.timing_log.clear()
        if dist is not None:
            dist.barrier()
        t = time.perf_counter()
        t1, tau, G = three_op_two_time(biexciton, t_axis, pulse, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4",
                                       tau_max=tau_max, dt=dt, options=dict(opts))
        if dist is not None:
            dist.barrier()       # the grid is complete when the slowest rank is
        return time.perf_counter() - t, G

    for _ in range(warmup):
        one_pass()
    t_begin = time.perf_counter()
    walls, logs = [], []
    for _ in range(steps):
        w, G = one_pass()
        walls.append(w)
        logs.append(list(eng.timing_log))
    t_end = time.perf_counter()
    eng.record_timings = False
    wall = float(np.mean(walls))
    main = [l for lg in logs for l in lg if l["kind"] == "main"]
    trunk = [l for lg in logs for l in lg if l["kind"] == "trunk"]
    NL = 16
    fl = flops_per_step(NL, chi)
    k_main = float(np.mean([l["step_ms"] for l in main]))
    k_trunk = float(np.mean([l["step_ms"] for l in trunk])) if trunk else 0.0
    steps_main = main[0]["traj_steps"]
    steps_trunk = trunk[0]["traj_steps"] if trunk else 0
    achieved = fl * steps_main / (k_main * 1e-3) / 1e12
    ref_steps = int(sum(round((t1 + tau_max) / dt) for t1 in t_axis))   # what the reference propagates (no forking)
    rec = {
        "workload": "cfg3: biexciton two-photon excitation + synthetic unitary PT (seed 1234), lindblad, "
                    "three_op_two_time G2(t,tau) %dx%d grid, dt=0.25 ps" % (n_t, n_t),
        "chi": chi, "NL": NL, "n_branches": n_t, "n_tau": n_t,
        "wall_ms": 1e3 * wall, "branch_kernel_ms": k_main, "trunk_kernel_ms": k_trunk,
        "opbuild_ms": float(np.mean([l["opbuild_ms"] for l in main])),
        "wall_over_kernels": 1e3 * wall / (k_main + k_trunk),
        "branch_kernel": main[0]["step_kernel"], "trunk_kernel": trunk[0]["step_kernel"] if trunk else None,
        "branch_steps": steps_main, "trunk_steps": steps_trunk,
        "roofline": {"bound": "tensor", "kernel": main[0]["step_kernel"] + " (branch launch)", "achieved": achieved,
                     "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma,
                     "achieved_trunk_and_branches": fl * (steps_main + steps_trunk) / ((k_main + k_trunk) * 1e-3) / 1e12,
                     "flops_per_trajectory_step": fl, "tile_T": main[0]["tile_T"], "cluster": main[0]["cluster"],
                     "n_tiles": main[0]["n_tiles"]},
        "trajectory_steps_per_s": (steps_main + steps_trunk) / wall,
        "reference_equivalent_steps": ref_steps, "reference_equivalent_steps_per_s": ref_steps / wall,
        "h2d_bytes_per_grid": int(2 * 16 * (2 * n_t)), "d2h_bytes_per_grid": int(G.nbytes * 2),
        "checks": {"max_abs_imag_tau0": float(np.abs(G[:, 0].imag).max()), "max_abs": float(np.abs(G).max())},
        "launches_per_grid": 2 * (len(main) + len(trunk)) // max(1, steps), "t_begin": t_begin, "t_end": t_end,
    }
    if cpu:
        # CPU restatement on MTO-free trajectories of the same shape (the MTO products are O(NL^2) per job)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_c
        from pyaceqd_b200.general_system import general_system as gs
        from pyaceqd_b200.jobs import FieldTable, Job
        prob = build_problem(system_op=["-4.0*|3><3|_4"], boson_op="1*(|1><1|_4 + |2><2|_4) + 2*|3><3|_4",
                             initial="|0><0|_4",
                             lindblad_ops=[["|0><1|_4", 0.01], ["|0><2|_4", 0.01], ["|1><3|_4", 0.01], ["|2><3|_4", 0.01]],
                             interaction_ops=[["|1><0|_4+|3><1|_4", "x"], ["|2><0|_4+|3><2|_4", "y"]],
                             output_ops=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)"])
        tt = np.arange(0.0, t_axis[-1] + tau_max, dt)
        px, py = gs.sample_pulses(tt, [pulse])
        tabs = {"x": FieldTable(0.0, dt, px), "y": FieldTable(0.0, dt, py)}
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else oracle_c.max_threads()
        idx = np.linspace(0, n_t - 1, max(cores, 16)).astype(int)
        jobs = [Job(0.0, float(t_axis[i] + tau_max), dt, tables=tabs) for i in idx]
        t = time.perf_counter()
        oracle_c.propagate_sweep(prob, pt, jobs, n_threads=cores)
        el = time.perf_counter() - t
        nst = sum(j.n_steps for j in jobs)
        rate = nst / el
        rec["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d of %d unforked trajectories, %d steps (%.1f s)" % (len(jobs), n_t, nst, el),
                               "wall_ms_extrapolated": 1e3 * ref_steps / rate,
                               "note": "CPU restatement (oracle/oracle_c.c, OpenMP), not ACE; the reference "
                                       "propagates every t1 from t=0 (no trunk sharing)"}
        rec["speedup_vs_cpu_grid_wall"] = rec["cpu_baseline"]["wall_ms_extrapolated"] / rec["wall_ms"]
    return rec


def run_cfg3(args):
    """``--workload cfg3``: the G2(t,tau) map as the bench line itself (grid wall time, branch-launch roofline)."""
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:      # the t axis shards over the ranks inside run_requests (opt-in); one all-gather of the rows
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        os.environ["ACEQD_DISTRIBUTED"] = "1"
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.5)
    rec = measure_g2_grid(local, n_t=args.n_t, chi=args.chi, steps=args.steps, warmup=args.warmup, dist=dist,
                          cpu=not args.no_cpu and rank == 0)
    sampler.t_begin, sampler.t_end = rec.pop("t_begin"), rec.pop("t_end")
    clocks = sampler.stop()
    from pyaceqd_b200.engine import default_engine
    if rank != 0:
        dist.destroy_process_group()
        default_engine(local).close()
        return
    line = {
        "metric": METRIC, "value": (rec["branch_steps"] + world * rec["trunk_steps"]) / (rec["wall_ms"] * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["wall_ms"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": rec["workload"], "chi": rec["chi"], "NL": rec["NL"], "n_branches": rec["n_branches"],
                   "n_tau": rec["n_tau"],
                   "l2": "per-row operators rebuilt and streamed every pass; PT slice (%.1f MB) is L2 resident by design"
                         % (9 * args.chi * args.chi * 16 / 1e6)},
        "clocks": clocks, "gpu_launches": rec["launches_per_grid"] * args.steps, "g2_grid_wall_ms": rec["wall_ms"],
        "e2e": {"value": rec["trajectory_steps_per_s"], "unit": UNIT, "h2d_bytes_per_step": rec["h2d_bytes_per_grid"],
                "d2h_bytes_per_step": rec["d2h_bytes_per_grid"],
                "api": "two_time.correlations.three_op_two_time(biexciton, ...) -> BatchExecutor -> aceqd_propagate_batch",
                "reference_equivalent_steps": rec["reference_equivalent_steps"],
                "reference_equivalent_steps_per_s": rec["reference_equivalent_steps_per_s"]},
        "roofline": dict(rec["roofline"], traffic=None, kernel_ms=rec["branch_kernel_ms"], trunk_kernel_ms=rec["trunk_kernel_ms"],
                         opbuild_ms=rec["opbuild_ms"], branch_steps=rec["branch_steps"], trunk_steps=rec["trunk_steps"]),
        "g2_checks": rec["checks"],
    }
    if "cpu_baseline" in rec:
        line["cpu_baseline"] = dict(rec["cpu_baseline"], g2_grid_wall_ms_extrapolated=rec["cpu_baseline"]["wall_ms_extrapolated"])
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    default_engine(local).close()


def measure_small_chi(eng, chi, n_area, n_det, n_steps, dt, steps=5, warmup=3, hbm_peak=None):
    """The bandwidth-bound regime of SURVEY 8d: the cfg2 sweep at a small bond dimension, device resident.  Reports the
    HBM bytes the pass moves (per-row operators written by the builder and read by the step kernel, drive tables,
    outputs) over its device time, against the measured HBM peak."""
    import torch
    prob, pt, tables_np = make_workload(chi, n_area, n_det, n_steps, dt)
    n_traj = tables_np.shape[0]
    plan = eng.plan_sweep(prob, pt, n_traj, n_steps, dt, 0.0, n_traj, n_steps, (0.0, dt))
    plan.batch.n_tables = 1
    tables_dev = torch.from_numpy(tables_np).cuda()
    out_dev = torch.empty((n_traj, n_steps + 1, prob.n_out), dtype=torch.complex128, device="cuda")
    k_ms, op_ms = [], []
    for i in range(warmup + steps):
        eng.run_sweep_device(prob, pt, plan, tables_dev.data_ptr(), out_dev.data_ptr())
        a, b = eng.last_timings()
        if i >= warmup:
            k_ms.append(a)
            op_ms.append(b)
    names = eng.last_kernels()
    rows = n_traj * (n_steps + 1)
    w_bytes = 8 * 4 * 16 + prob.n_out * 4 * 16                  # W [8][4] + OV [n_out][4] complex per trajectory-row
    traffic = rows * (16.0 * prob.n_out) + tables_np.nbytes      # outputs + drive tables: what the path must move
    if float(np.mean(op_ms)) > 0.0:                              # a separate operator builder ran:
        traffic += 2.0 * rows * w_bytes                          # operators written by it, read back by the step kernel
    ms = float(np.mean(k_ms)) + float(np.mean(op_ms))
    rec = {"workload": "cfg2 at chi=%d: %dx%d sweep, %d steps" % (chi, n_area, n_det, n_steps), "chi": chi,
           "step_kernel": names["step"], "opbuild_kernel": names["opbuild"],
           "kernel_ms": float(np.mean(k_ms)), "opbuild_ms": float(np.mean(op_ms)), "pass_ms": ms,
           "trajectory_steps_per_s": n_traj * n_steps / (ms * 1e-3),
           "hbm_bytes_per_pass": traffic, "algorithmic_bytes_per_pass": rows * 16.0 * prob.n_out + tables_np.nbytes,
           "roofline": {"bound": "hbm", "achieved": traffic / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (traffic / (ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None},
           "dmma_flops_per_trajectory_step": flops_per_step(4, chi)}
    return rec


def measure_strong(eng, world, rank, dist, n_t=96, chi=256, tb=24.0, dt=0.1):
    """Strong scaling of ONE cfg5-shaped sweep (five-level dark model NL=25, chi=256, triangular
    (t1, t2) sweep with three multi-time operators per run, timebin/twophoton_new.py:515-557): every rank holds the
    same job list, `run_jobs_sharded` gives each a contiguous, step-balanced block and all-gathers the kept rows.
    Timed on all ranks (max); with world > 1 rank 0 also runs the whole sweep alone for the efficiency."""
    import torch
    from pyaceqd_b200.distributed import run_jobs_sharded
    from pyaceqd_b200.jobs import FieldTable, Job
    from pyaceqd_b200.problem import build_problem
    from pyaceqd_b200.process_tensor import synthetic_pt
    from pyaceqd_b200.pulses import ChirpedPulse
    prob = build_problem(
        system_op=["-4.0*|4><4|_5", "-0.1*|3><3|_5"], boson_op="1*(|1><1|_5 + |2><2|_5 + |3><3|_5) + 2*|4><4|_5",
        initial="|0><0|_5", lindblad_ops=[["|0><1|_5", 0.01], ["|0><2|_5", 0.01], ["|1><4|_5", 0.01], ["|2><4|_5", 0.01]],
        interaction_ops=[["|1><0|_5", "x"], ["|4><1|_5", "x"], ["|3><0|_5", "y"], ["|4><3|_5", "y"]],
        output_ops=["|0><1|_5", "|0><1|_5*|1><4|_5"])
    pt = synthetic_pt(chi, len(prob.cls_keys), dt=dt, seed=1234, kind="unitary", scale=0.999)
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=4.0, e0=5.0, polar_x=1.0)
    t1 = np.round(np.linspace(0.0, tb, n_t), 1)
    tt = np.arange(0.0, 2 * tb + 1.0, dt)
    f = p.get_total(tt)
    tabs = {"x": FieldTable(0.0, dt, np.round(f.real, 8) + 1j * np.round(f.imag, 8)),
            "y": FieldTable(0.0, dt, np.zeros_like(f))}
    jobs = []
    for i in range(n_t):
        for j in range(i, n_t):
            mt = prob.parse_mtos([{"operator": "|4><1|_5", "applyFrom": "_right", "time": float(t1[i])},
                                  {"operator": "|1><0|_5", "applyFrom": "_right", "time": float(t1[j])},
                                  {"operator": "|1><4|_5", "applyFrom": "_left", "time": float(t1[i] + tb)}])
            jobs.append(Job(0.0, float(t1[j] + tb), dt, tables=tabs, mtos=mt, tail_rows=1))
    steps_total = sum(j.n_steps for j in jobs)

    def timed(sharded):
        if dist is not None and sharded:     # (the one-GPU reference leg runs on rank 0 alone: no collective in it)
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = run_jobs_sharded(eng, prob, pt, jobs) if sharded else eng.run_jobs(prob, pt, jobs)
        torch.cuda.synchronize()
        el = time.perf_counter() - t
        if dist is not None and sharded:
            te = torch.tensor([el], dtype=torch.float64, device="cuda")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            el = float(te.item())
        return el, out

    timed(world > 1)                                  # warm-up (handles, workspace)
    t_n, out_n = timed(world > 1)
    rec = {"workload": "cfg5-shaped: five-level NL=25, synthetic PT chi=%d, triangular (t1,t2) sweep %dx%d/2 = %d runs with "
                       "3 multi-time operators, tb=%g ps, dt=%g ps (reference-equivalent %d trajectory-steps)"
                       % (chi, n_t, n_t, len(jobs), tb, dt, steps_total),
           "n_gpus": world, "wall_s": t_n, "reference_equivalent_steps_per_s": steps_total / t_n,
           "kernel": eng.last_kernels()["step"], "checksum": float(np.abs(np.concatenate([o[:, -1] for o in out_n])).sum())}
    if world > 1:
        if rank == 0:
            t_1, out_1 = timed(False)
            rec["wall_s_one_gpu_same_box"] = t_1
            rec["efficiency_vs_one_gpu"] = t_1 / (world * t_n)
            rec["max_abs_diff_sharded_vs_one_gpu"] = float(max(np.abs(a - b).max() for a, b in zip(out_1, out_n)))
        dist.barrier()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--n-area", type=int, default=64)
    ap.add_argument("--n-det", type=int, default=64)
    ap.add_argument("--n-steps", type=int, default=400)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--n-t", type=int, default=256, help="cfg3: points of the t and tau axes")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records (g2_grid = cfg3 map, small_chi = bandwidth regime, strong = sharded cfg5-shaped sweep)")
    args = ap.parse_args()
    if args.workload == "cfg3" and args.impl == "ours":
        return run_cfg3(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dt = 0.1
    n_traj = args.n_area * args.n_det
    config = {"workload": "cfg2: two-level QD + synthetic PT (seed 1234), %dx%d pulse-area x detuning sweep, "
                          "%d steps of dt=0.1 ps" % (args.n_area, args.n_det, args.n_steps),
              "chi": args.chi, "NL": 4, "n_traj_per_gpu": n_traj, "n_steps": args.n_steps,
              "l2": "inputs larger than L2: %.2f GB of per-step operators rebuilt and streamed every pass"
                    % (n_traj * (args.n_steps + 1) * (512 + 256) / 1e9)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        prob, pt, tables = make_workload(args.chi, args.n_area, args.n_det, args.n_steps, dt)
        vals, times = [], []
        sample, cores = "", 0
        for i in range(args.warmup + args.steps):
            t_ = time.perf_counter()
            r, sample, cores = cpu_rate(prob, pt, tables, args.n_steps, dt,
                                        seconds=max(2.0, 60.0 / (args.warmup + args.steps)))
            if i >= args.warmup:
                vals.append(r)
                times.append(time.perf_counter() - t_)
        v = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement (oracle/oracle_c.c), not ACE: the reference's own path is the external ACE binary"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from pyaceqd_b200.engine import Engine

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a real (non-default) stream: handle 0 would make the library create its own stream and the
    # torch events below would not see the kernels
    # high priority: the library builds the later waves' operators on a low-priority side stream that
    # should only fill the SMs the first (partial) wave of persistent CTAs leaves idle
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = Engine(local, stream=stream.cuda_stream)
    prob, pt, tables_np = make_workload(args.chi, args.n_area, args.n_det, args.n_steps, dt, rank=rank)
    n_out, NL = prob.n_out, prob.NL
    tables_pin = eng.pinned_empty(tables_np.shape, np.complex128)
    tables_pin[...] = tables_np
    plan = eng.plan_sweep(prob, pt, n_traj, args.n_steps, dt, 0.0, n_traj, args.n_steps, (0.0, dt),
                          tile_T=args.tile or None)
    plan.batch.n_tables = 1
    tables_dev = torch.from_numpy(tables_np).cuda()
    # two output buffers: with several ranks the all-gather of pass k runs on a side stream while pass k+1 computes
    out_bufs = [torch.empty((n_traj, args.n_steps + 1, n_out), dtype=torch.complex128, device="cuda")
                for _ in range(2 if world > 1 else 1)]
    out_dev = out_bufs[0]
    gathered = [torch.empty((world,) + tuple(out_dev.shape), dtype=torch.complex128, device="cuda") for _ in out_bufs] \
        if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None
    ev_done = [torch.cuda.Event() for _ in out_bufs]      # pass written into buffer b
    ev_gath = [torch.cuda.Event() for _ in out_bufs]      # buffer b gathered (may be overwritten)
    pass_no = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        b = pass_no[0] % len(out_bufs)
        pass_no[0] += 1
        if world > 1 and pass_no[0] > len(out_bufs):
            stream.wait_event(ev_gath[b])                 # the gather that read this buffer two passes ago is done
        eng.run_sweep_device(prob, pt, plan, tables_dev.data_ptr(), out_bufs[b].data_ptr())
        if world > 1:
            ev_done[b].record(stream)
            with torch.cuda.stream(side):
                side.wait_event(ev_done[b])
                dist.all_gather_into_tensor(gathered[b], out_bufs[b])
                ev_gath[b].record(side)

    def join_gathers():
        if world > 1:
            stream.wait_stream(side)                      # the last gathers end inside the timed region

    peak_dmma = eng.fp64_peak("dmma", 20000)
    peak_dfma = eng.fp64_peak("dfma", 20000)

    dbg = os.environ.get("BENCH_DEBUG")
    tdbg = time.perf_counter()

    def mark(msg):
        nonlocal tdbg
        if dbg and rank == 0:
            torch.cuda.synchronize()
            now = time.perf_counter()
            sys.stderr.write("[bench] %-28s %.1f ms\n" % (msg, 1e3 * (now - tdbg)))
            tdbg = now

    mark("setup")
    sampler = ClockSampler(local)   # started before warm-up: the first nvidia-smi start on a box can stall
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    mark("sampler start")
    for _ in range(args.warmup):
        one_pass()
        mark("warmup pass")
    join_gathers()
    barrier()
    sampler.t_begin = time.perf_counter()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    k_ms, op_ms = [], []
    for _ in range(args.steps):
        one_pass()
    join_gathers()
    e1.record(stream)
    barrier()
    sampler.t_end = time.perf_counter()
    mark("timed passes")
    launches = eng.launch_count() - n0
    ms_total = e0.elapsed_time(e1)
    k_last, op_last = eng.last_timings()
    # per-launch step-kernel durations: re-run K passes reading the library's own events
    for _ in range(args.steps):
        eng.run_sweep_device(prob, pt, plan, tables_dev.data_ptr(), out_dev.data_ptr())
        a, b = eng.last_timings()
        k_ms.append(a)
        op_ms.append(b)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    units = float(world) * n_traj * args.n_steps * args.steps
    value = units / (ms_total * 1e-3)

    # ---- end to end: host (pinned) buffers through the public API
    # (the sweep is planned inside every call: nothing of the host path is hoisted out of the timed loop)
    for _ in range(2):
        res = eng.run_sweep(prob, pt, tables_pin, (0.0, dt), 0.0, args.n_steps, dt, copy=False, tile_T=args.tile or None)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = eng.run_sweep(prob, pt, tables_pin, (0.0, dt), 0.0, args.n_steps, dt, copy=False, tile_T=args.tile or None)
        final_x = float(res[-1, -1, 1].real)   # read a result on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units / float(te.item())
    # consistency of both legs (same inputs -> same numbers)
    dev_host = out_bufs[(pass_no[0] - 1) % len(out_bufs)].cpu().numpy()
    leg_diff = float(np.abs(dev_host - res).max())

    if rank == 0:
        k_avg = float(np.mean(k_ms))
        fl = flops_per_step(NL, args.chi) * n_traj * args.n_steps
        achieved = fl / (k_avg * 1e-3) / 1e12
        algo_bytes = n_traj * (args.n_steps + 1) * (768.0 + 16.0 * n_out)
        traffic, traffic_src = None, None
        try:    # DRAM bytes of one step-kernel launch from the committed ncu --set full capture of this exact config
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                ent = json.load(fh).get("cfg2_chi%d_%dx%dx%d" % (args.chi, args.n_area, args.n_det, args.n_steps))
            if ent:
                traffic, traffic_src = float(ent["dram_bytes_read"]) + float(ent["dram_bytes_write"]), ent["source"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tables_pin.nbytes),
                    "d2h_bytes_per_step": int(res.nbytes), "ms_per_step": 1e3 * float(te.item()) / args.steps,
                    "api": "Engine.run_sweep -> plan_sweep + aceqd_propagate_batch (host pinned buffers; planned inside every call)"},
            "roofline": {"bound": "tensor", "kernel": "k_step_dmma", "achieved": achieved, "peak": peak_dmma,
                         "unit": "TFLOP/s", "frac": achieved / peak_dmma, "traffic": traffic,
                         "traffic_note": "DRAM bytes (read+write) of one step-kernel launch of this exact config from "
                                         "the ncu --set full capture %s (includes the bond states handed between CTAs by "
                                         "the segment schedule); algorithmic bytes (per-row operators + outputs) = %.4g"
                                         % (traffic_src, algo_bytes),
                         "peak_source": "FP64 DMMA.8x8x4 register-resident micro-benchmark (aceqd_fp64_peak) measured in this run; "
                                        "MEASURED_PEAKS.json holds no FP64 figure; nominal B200 FP64 ~40 TFLOP/s",
                         "dfma_peak": peak_dfma, "kernel_ms": k_avg, "opbuild_ms": float(np.mean(op_ms)),
                         "flops_per_trajectory_step": flops_per_step(NL, args.chi),
                         "tile_T": int(plan.batch.tile_T), "n_tiles": int(plan.batch.n_tiles)},
            "legs_max_abs_diff": leg_diff, "final_x_last_traj": final_x,
        }
    # ---- sub-records: the north-star map, the bandwidth regime and a strong-scaling sweep (same JSON line)
    extras = {}
    if not args.no_extras:
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_peak = float(json.load(fh)["hbm_gbs"])
        except (OSError, ValueError, KeyError):
            hbm_peak = 6555.0       # fallback of /opt/skills/guides/B200_PROFILING.md
        try:
            extras["strong"] = measure_strong(eng, world, rank, dist if world > 1 else None)
        except Exception as exc:    # noqa: BLE001 - a sub-record must not take the headline down
            extras["strong"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if rank == 0:
            try:
                extras["g2_grid"] = measure_g2_grid(local, cpu=not args.no_cpu and world == 1, peak_dmma=peak_dmma)
                extras["g2_grid"].pop("t_begin", None), extras["g2_grid"].pop("t_end", None)
            except Exception as exc:    # noqa: BLE001
                extras["g2_grid"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            try:
                extras["small_chi"] = measure_small_chi(eng, 16, args.n_area, args.n_det, args.n_steps, dt, hbm_peak=hbm_peak)
                extras["small_chi"]["roofline"]["peak_source"] = "MEASURED_PEAKS.json hbm_gbs" if hbm_peak != 6555.0 else \
                    "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"
            except Exception as exc:    # noqa: BLE001
                extras["small_chi"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        line.update(extras)
        # the CPU baseline runs LAST: its OpenMP team keeps spinning for a while and slows the host side of whatever
        # is measured next (the G2 map's wall time doubled when it ran first)
        if not args.no_cpu and world == 1:
            r, sample, cores = cpu_rate(prob, pt, tables_np, args.n_steps, dt, seconds=args.cpu_seconds)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "note": "CPU restatement (oracle/oracle_c.c, OpenMP: -march=x86-64-v3, one trajectory "
                                            "per thread, PT walked per trajectory), not ACE"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
