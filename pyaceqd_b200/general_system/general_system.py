"""``system_ace_stream`` -- the drop-in boundary.

Same signature, argument meaning, return layout and error behaviour as the reference's
``pyaceqd/general_system/general_system.py:128-360``, but instead of writing an ACE parameter
file + pulse files and spawning the ``ACE`` binary (``:227-343``) it builds a numeric
:class:`~pyaceqd_b200.problem.Problem`, samples the drives exactly like the pulse files
(grid ``np.arange(t_start, t_end, dt)`` ``:213``, 8 decimals ``:69-70``) and runs the CUDA engine
in process.  Inside a :class:`~pyaceqd_b200.batch.BatchExecutor` the call is deferred so that a
whole sweep becomes one GPU batch.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.jobs import FieldTable, Job
from pyaceqd_b200.problem import Problem, build_problem
from pyaceqd_b200.process_tensor import ProcessTensor

hbar = constants.hbar  # meV*ps
temp_dir = constants.temp_dir

_capture = threading.local()          # set by BatchExecutor: deferred execution
_problem_cache: Dict[tuple, Problem] = {}
_pt_cache: Dict[str, ProcessTensor] = {}
_table_cache: Dict[tuple, Dict[str, FieldTable]] = {}
_cache_lock = threading.Lock()


def sanity_checks(system_op, phonons, boson_op, initial, interaction_ops, verbose):
    """Reference ``general_system.py:17-27`` (prints + exit on a missing boson operator)."""
    if system_op is None and verbose:
        print("System operator not supplied, assuming TLS")
    if phonons and boson_op is None:
        print("using phonons, but boson operator not specified")
        exit(1)
    if initial is None and verbose:
        print("No initial state specified")
    if interaction_ops is None and verbose:
        print("No interaction hamiltonian ")


def check_multitime(multitime_op, verbose):
    """Normalise one multitime dict in place (reference ``general_system.py:29-53``)."""
    if verbose:
        print("multitime operator: {}".format(multitime_op))
    if multitime_op is None:
        return
    if "operator" not in multitime_op or "time" not in multitime_op:
        print("supply 'operator' and 'time' for multitime")
        exit(0)
    multitime_op.setdefault("applyFrom", "")       # "": A rho A^+
    multitime_op.setdefault("applyBefore", "false")  # after `time`: visible at time+dt
    if multitime_op["applyFrom"] not in ("_left", "_right", ""):
        print(multitime_op)
        print('give "_left" or "_right" or "" for multitime')
        exit(0)


def _quantise(v: np.ndarray) -> np.ndarray:
    """The ``%.8f`` text round trip of the pulse files (``general_system.py:69-70``)."""
    return np.round(np.real(v), 8) + 1j * np.round(np.imag(v), 8)


def sample_pulses(t, pulses, abs_only=False) -> Tuple[np.ndarray, np.ndarray]:
    """x / y drive samples, ``sum_p polar * p.get_total(t)`` (reference ``generate_pulsefiles``,
    ``general_system.py:55-71``), quantised like the pulse file."""
    px = np.zeros_like(t, dtype=complex)
    py = np.zeros_like(t, dtype=complex)
    for p in pulses:
        f = p.get_total(t)
        if abs_only:
            f = np.abs(f)
        px = px + p.polar_x * f
        py = py + p.polar_y * f
    return _quantise(px), _quantise(py)


def sample_rf(t, pulses, firstonly=False):
    """Rotating-frame tables (reference ``generate_rf_file``, ``general_system.py:73-102``): the rf
    table is the instantaneous frequency of the first pulse; every carrier is shifted by the
    first pulse's ``e_start`` and all chirps are zeroed before the x / y tables are regenerated."""
    if len(pulses) > 1:
        print("Warning: more than one pulse supplied, only the first one is used for rf")
        print("Note that also, chirping more than the first pulse is not supported")
    rf = _quantise(np.asarray(pulses[0].get_frequency(t), dtype=complex) * np.ones_like(t))
    shifted = [p.copy() for p in pulses]
    e0, _ = shifted[0].get_energy()
    for p in shifted:
        e, _ = p.get_energy()
        p.set_energy(e - e0, 0)
    px, py = sample_pulses(t, [shifted[0]] if firstonly else shifted)
    return rf, px, py


def generate_pulsefiles(t, pulses, temp_dir, system_prefix, suffix, abs_only=False):
    """The reference's pulse-file writer under its own name (``general_system.py:55-71``): ``t Re Im`` rows with 8 decimals
    for the x and y components; returns the two file names.  (The in-process path never writes them: it hands
    :func:`sample_pulses` to the operator builder; scripts that call this function directly get the reference's files.)"""
    from pyaceqd_b200.tools import export_csv
    fx = temp_dir + "{}_pulse_x_{}.dat".format(system_prefix, suffix)
    fy = temp_dir + "{}_pulse_y_{}.dat".format(system_prefix, suffix)
    px, py = sample_pulses(np.asarray(t), pulses, abs_only=abs_only)
    export_csv(fx, t, px.real, px.imag, precision=8, delimit=' ')
    export_csv(fy, t, py.real, py.imag, precision=8, delimit=' ')
    return fx, fy


def generate_rf_file(t, pulses, temp_dir, system_prefix, suffix, firstonly=False):
    """Rotating-frame file of the first pulse's instantaneous frequency, and the x / y pulse files regenerated in that
    frame (reference ``general_system.py:73-102``); returns the rf file name."""
    from pyaceqd_b200.tools import export_csv
    rf_file = temp_dir + "{}_rf_{}.dat".format(system_prefix, suffix)
    rf, px, py = sample_rf(np.asarray(t), pulses, firstonly=firstonly)
    export_csv(rf_file, t, rf.real, rf.imag, precision=8, delimit=' ')
    export_csv(temp_dir + "{}_pulse_x_{}.dat".format(system_prefix, suffix), t, px.real, px.imag, precision=8, delimit=' ')
    export_csv(temp_dir + "{}_pulse_y_{}.dat".format(system_prefix, suffix), t, py.real, py.imag, precision=8, delimit=' ')
    return rf_file


def read_pulse_file(path: str) -> FieldTable:
    """ACE pulse-file reader: columns ``t Re Im`` (``general_system.py:69-70``), cached by mtime.  Names registered by
    ``PulseGenerator.generate_pulsefiles(in_memory=True)`` resolve without touching the disk."""
    from pyaceqd_b200.pulsegenerator import memory_file
    mem = memory_file(path)
    if mem is not None:
        t, values = mem
        return FieldTable(float(t[0]), float(t[1] - t[0]) if len(t) > 1 else 1.0, values)
    key = ("file", os.path.abspath(path), os.path.getmtime(path))
    with _cache_lock:
        if key in _table_cache:
            return _table_cache[key]["f"]
    data = np.loadtxt(path, ndmin=2)
    t = data[:, 0]
    dt = float(t[1] - t[0]) if len(t) > 1 else 1.0
    tab = FieldTable(float(t[0]), dt, data[:, 1] + 1j * data[:, 2])
    with _cache_lock:
        _table_cache[key] = {"f": tab}
    return tab


def read_result(data, n):
    """``[t, Re o1, Im o1, ...]`` rows -> ``[(1+n), n_t]`` complex (reference ``:104-110``)."""
    t = data[:, 0]
    result = np.empty([1 + n, len(t)], dtype=complex)
    result[0] = t
    for i in range(n):
        result[i + 1] = data[:, 2 * i + 1] + 1j * data[:, 2 * i + 2]
    return result


def read_result_1d(data):
    """As :func:`read_result` with the number of outputs taken from the column count (reference ``:112-119``)."""
    return read_result(data, int((data.shape[1] - 1) / 2))


def read_hamiltonian(data):
    """``print_H`` output rows ``[_, Re h_0, Im h_0, ...]`` -> complex ``n x n`` matrix, column by column (reference
    ``:121-126``)."""
    n = data.shape[0]
    result = np.empty([n, n], dtype=complex)
    for i in range(n):
        result[:, i] = data[:, 2 * i + 1] + 1j * data[:, 2 * i + 2]
    return result


@dataclass
class Request:
    """One deferred ``system_ace_stream`` call."""
    problem: Problem
    pt: Optional[ProcessTensor]
    job: Job
    pulse_key: tuple          # identifies the drive (pulse objects / files) for table sharing
    table_maker: object       # callable(t_end) -> tables dict, used when tables are unified
    calc_dynmap: bool = False
    result: Optional[np.ndarray] = None
    sampled: bool = True      # drives sampled from pulse objects on np.arange(t_start, t_end, dt) (not read from files)


def _problem_for(**kw) -> Problem:
    key = (tuple(kw.get("system_op") or ()), kw.get("boson_op"), kw.get("initial"),
           tuple((o, float(r)) for o, r in (kw.get("lindblad_ops") or ())),
           tuple((o, p) for o, p in (kw.get("interaction_ops") or ())),
           tuple(kw.get("output_ops") or ()), kw.get("rf_op"),
           None if kw.get("rho0") is None else np.asarray(kw["rho0"]).tobytes(), kw.get("dict_zero"))
    with _cache_lock:
        if key not in _problem_cache:
            _problem_cache[key] = build_problem(
                system_op=kw.get("system_op"), boson_op=kw.get("boson_op"), initial=kw.get("initial"),
                lindblad_ops=kw.get("lindblad_ops"), interaction_ops=kw.get("interaction_ops"),
                output_ops=kw.get("output_ops") or (), rf_op=kw.get("rf_op"), rho0=kw.get("rho0"),
                dict_zero=10.0 ** (-int(kw.get("dict_zero") or 16)))
        return _problem_cache[key]


def resolve_pt(pt_file: str, *, boson_op, dt, t_mem, ae, temperature, threshold, factor_ah, use_infinite,
               boson_e_max, J_file, verbose, problem: Problem) -> ProcessTensor:
    """Load (or build and cache) the phonon PT named like the reference does (``:146-151``)."""
    with _cache_lock:
        if pt_file in _pt_cache:
            return _pt_cache[pt_file]
    if os.path.exists(pt_file):
        pt = ProcessTensor.load(pt_file)
        if verbose:
            print("using pt_file " + pt_file)
    else:
        if os.path.exists(pt_file + "_initial"):
            raise NotImplementedError(
                "{}_initial is an ACE-format process tensor; its binary layout is undocumented "
                "(SURVEY App. E, R2). Rebuild it with this engine (delete the ACE files).".format(pt_file))
        print("{} not found. Calculating...".format(pt_file))
        from pyaceqd_b200.pt_builder import build_pt_from_spectral_density_file, build_qd_phonon_pt
        if J_file is not None:      # Boson_J_from_file (reference :178-179): tabulated spectral density
            pt = build_pt_from_spectral_density_file(J_file, problem.meta["coupling_diag"], float(dt), float(t_mem),
                                                     float(temperature), threshold=10.0 ** (-int(threshold)),
                                                     e_max=float(boson_e_max), verbose=verbose)
        else:
            pt = build_qd_phonon_pt(coupling_diag=problem.meta["coupling_diag"], dt=float(dt), t_mem=float(t_mem),
                                    a_e=float(ae), a_h=None if factor_ah is None else float(ae) / float(factor_ah),
                                    temperature=float(temperature), threshold=10.0 ** (-int(threshold)),
                                    e_max=float(boson_e_max), use_infinite=bool(use_infinite), verbose=verbose)
        try:
            pt.save(pt_file)
        except OSError:
            pass
    with _cache_lock:
        _pt_cache[pt_file] = pt
    return pt


def system_ace_stream(t_start, t_end, *pulses, dt=0.01, phonons=False, t_mem=20.48, ae=3.0, temperature=1,
                      verbose=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None,
                      pulse_file_x=None, pulse_file_y=None, system_prefix="", threshold="10",
                      threshold_ratio="0.3", buffer_blocksize="-1", dict_zero="16", precision="12",
                      boson_e_max=7, system_op=None, boson_op=None, initial=None, lindblad_ops=None,
                      interaction_ops=None, output_ops=[], prepare_only=False, LO_params=None,
                      dressedstates=False, rf_op=None, rf_file=None, firstonly=False, J_to_file=None,
                      J_file=None, factor_ah=None, use_infinite=False, print_H=False, calc_dynmap=False,
                      rho0=None, get_M_t=None):
    """In-process replacement of the ACE round trip; see the module docstring.

    Returns ``ndarray[(1+len(output_ops)), n_t]`` complex with row 0 = time
    (reference ``:343,360``); ``(result, dm[n_t, NL, NL])`` with ``calc_dynmap`` (``:358-359``);
    the single-step propagator matrix with ``get_M_t`` (``:325-327``).
    """
    sanity_checks(system_op=system_op, phonons=phonons, boson_op=boson_op, initial=initial,
                  interaction_ops=interaction_ops, verbose=verbose)
    if multitime_op is not None:
        if isinstance(multitime_op, dict):
            multitime_op = [multitime_op]
        for _mto in multitime_op:
            check_multitime(multitime_op=_mto, verbose=verbose)
    if LO_params is not None:
        raise NotImplementedError("add_single_mode (LO_params) needs a non-diagonal environment; out of scope")
    if dressedstates or print_H:
        raise NotImplementedError("timedep_eigenstates / print_H are ACE diagnostics binaries; out of scope")
    if prepare_only:
        # like the reference (:292-296): write the parameter file + pulse files and return dummies.
        # The files are what `ACE <file>` (scripts/ACE -> pyaceqd_b200.ace_cli) consumes.
        from pyaceqd_b200.ace_cli import write_param_file
        from pyaceqd_b200.tools import export_csv
        stem = temp_dir + "{}_{}".format(system_prefix, suffix)
        t = np.arange(t_start, t_end, step=dt / 1)
        px_file, py_file, rf_out = pulse_file_x, pulse_file_y, rf_file
        if rf_op is not None and rf_file is None:
            rf, px, py = sample_rf(t, pulses, firstonly=firstonly)
            rf_out = temp_dir + "{}_rf_{}.dat".format(system_prefix, suffix)       # reference :77
            export_csv(rf_out, t, rf.real, rf.imag, precision=8, delimit=' ')
            px_file = None
        elif pulse_file_x is None:
            px, py = sample_pulses(t, [pulses[0]] if firstonly else pulses)
        if px_file is None:
            px_file = temp_dir + "{}_pulse_x_{}.dat".format(system_prefix, suffix)     # reference :56-57
            py_file = temp_dir + "{}_pulse_y_{}.dat".format(system_prefix, suffix)
            export_csv(px_file, t, px.real, px.imag, precision=8, delimit=' ')
            export_csv(py_file, t, py.real, py.imag, precision=8, delimit=' ')
        if phonons and pt_file is None:
            pt_file = "{}_{}nm_{}k_th{}_{}dt{}.pt{}".format(system_prefix, ae, temperature, threshold,
                                                          "" if use_infinite else "tmem{}_".format(t_mem), dt,
                                                          "" if use_infinite else "r")
        write_param_file(stem + ".param", dt=dt, t_start=t_start, t_end=t_end, dict_zero=dict_zero,
                         precision=precision, pt_file=pt_file if phonons else None, initial=initial,
                         system_op=system_op, rf_op=rf_op, rf_file=rf_out, lindblad_ops=lindblad_ops,
                         interaction_ops=interaction_ops, pulse_file_x=px_file, pulse_file_y=py_file,
                         multitime_op=multitime_op, output_ops=output_ops, out_file=stem + ".out")
        print("prepared file {}, exiting.".format(stem + ".param"))
        return [np.array([0, 0]) for _ in range(1 + len(output_ops))]

    if J_to_file is not None and phonons:      # Boson_J_print <file> 0 15 2000 (reference :186-187)
        from pyaceqd_b200.pt_builder import write_spectral_density
        write_spectral_density(J_to_file, a_e=float(ae), a_h=None if factor_ah is None else float(ae) / float(factor_ah))
    problem = _problem_for(system_op=system_op, boson_op=boson_op if phonons else None, initial=initial,
                           lindblad_ops=lindblad_ops, interaction_ops=interaction_ops, output_ops=output_ops,
                           rf_op=rf_op, rho0=rho0, dict_zero=dict_zero)

    pt = None
    if phonons:
        if pt_file is None:
            pt_file = "{}_{}nm_{}k_th{}_tmem{}_dt{}.ptr".format(system_prefix, ae, temperature, threshold, t_mem, dt)
            if J_file is not None:
                pt_file = "{}_{}_{}k_th{}_tmem{}_dt{}.ptr".format(system_prefix, os.path.splitext(J_file)[0],
                                                                 temperature, threshold, t_mem, dt)
            if use_infinite:
                # the reference name omits ae (cache-collision hazard); keep ae in ours
                # ... and t_mem: the builder truncates the memory kernel there in the infinite mode too, so a
                # different t_mem must not silently reuse the cached tensor
                pt_file = "{}_{}nm_{}k_th{}_tmem{}_dt{}.pt".format(system_prefix, ae, temperature, threshold, t_mem, dt)
        pt = resolve_pt(pt_file, boson_op=boson_op, dt=dt, t_mem=t_mem, ae=ae, temperature=temperature,
                        threshold=threshold, factor_ah=factor_ah, use_infinite=use_infinite,
                        boson_e_max=boson_e_max, J_file=J_file, verbose=verbose, problem=problem)

    # ---- drive tables: the content of the pulse files
    def make_tables(te):
        t = np.arange(t_start, te, step=dt / 1)   # reference :213
        tabs: Dict[str, FieldTable] = {}
        if rf_op is not None and rf_file is None:
            rf, px, py = sample_rf(t, pulses, firstonly=firstonly)
            tabs["rf"] = FieldTable(float(t_start), float(dt), rf)
            tabs["x"] = FieldTable(float(t_start), float(dt), px)
            tabs["y"] = FieldTable(float(t_start), float(dt), py)
            return tabs
        if rf_op is not None:
            tabs["rf"] = read_pulse_file(rf_file)
        if pulse_file_x is not None:
            tabs["x"] = read_pulse_file(pulse_file_x)
            if pulse_file_y is not None:
                tabs["y"] = read_pulse_file(pulse_file_y)
        else:
            px, py = sample_pulses(t, [pulses[0]] if firstonly else pulses)
            tabs["x"] = FieldTable(float(t_start), float(dt), px)
            tabs["y"] = FieldTable(float(t_start), float(dt), py)
        return tabs

    if interaction_ops is not None:
        for _op in interaction_ops:
            if _op[1] == "y" and pulse_file_x is not None and pulse_file_y is None:
                print("Pulse file y not given")
                exit(1)
    pulse_key = (tuple(id(p) for p in pulses), pulse_file_x, pulse_file_y, rf_file, float(t_start), float(dt),
                 bool(firstonly), rf_op)

    if get_M_t is not None:
        # fprop.update(t, dt); fprop.M  (:324-327): propagator of one full step starting at t
        from pyaceqd_b200.engine import default_engine
        job = Job(float(t_start), float(t_end), float(dt), tables=make_tables(t_end))
        L = problem.L0.copy()
        for k, pol in enumerate(problem.field_pol):
            tab = job.tables.get(pol)
            if tab is None:
                continue
            x = (get_M_t + 0.5 * dt - tab.t0) / tab.dt
            f = np.interp(x, np.arange(len(tab.values)), tab.values.real) + \
                1j * np.interp(x, np.arange(len(tab.values)), tab.values.imag)
            L = L + f * problem.LA[k] + np.conj(f) * problem.LB[k]
        return default_engine().expm(L * dt)[0]

    sink = getattr(_capture, "sink", None)
    # deferred calls sample their drives once per sweep in run_requests (shared, longest window)
    job = Job(float(t_start), float(t_end), float(dt),
              tables={} if sink is not None else make_tables(t_end),
              mtos=problem.parse_mtos(multitime_op))
    req = Request(problem=problem, pt=pt, job=job, pulse_key=pulse_key, table_maker=make_tables,
                  calc_dynmap=calc_dynmap,
                  sampled=(rf_op is not None and rf_file is None) or (rf_op is None and pulse_file_x is None))
    if sink is not None:
        sink.append(req)
        return req      # BatchExecutor resolves it
    return run_requests([req])[0]


def _finish(req: Request, out: np.ndarray):
    res = np.empty((1 + out.shape[0], out.shape[1]), dtype=complex)
    res[0] = req.job.times()[-out.shape[1]:]   # deferred jobs may keep only their last rows
    res[1:] = out
    return res


def run_requests(reqs: List[Request], distributed: Optional[bool] = None, tail_reduce=None):
    """Execute deferred requests: one GPU batch per (problem, PT, dt, drive-table grid) group.

    Requests whose drives are sampled from different start times live on different table grids and run as
    separate batches (the reference handles them independently as well).  Inside a ``torch.distributed`` job the
    sweep is sharded over the ranks ONLY when asked for: ``distributed=True`` here, ``BatchExecutor(distributed=True)``
    or ``ACEQD_DISTRIBUTED=1`` -- every rank must then submit the same job list in the same order (checked).
    ``tail_reduce = (pairs, spacing)``: every request returns the tau integral of its kept rows instead of the rows
    (``Engine.run_jobs``), reduced on the device."""
    from pyaceqd_b200.engine import default_engine
    eng = default_engine()
    if distributed is None:
        distributed = os.environ.get("ACEQD_DISTRIBUTED", "0") == "1"
    results = [None] * len(reqs)
    groups: Dict[tuple, List[int]] = {}
    for i, r in enumerate(reqs):
        # pulse_key[4] is the t_start the drive tables are sampled from: one table grid per batch
        groups.setdefault((id(r.problem), id(r.pt), r.job.dt, round(r.job.t_start / r.job.dt, 6)), []).append(i)
    for _, idx in groups.items():
        prob, pt = reqs[idx[0]].problem, reqs[idx[0]].pt
        # unify drive tables of requests that sample the same pulses from the same t_start
        by_key: Dict[tuple, List[int]] = {}
        for i in idx:
            by_key.setdefault(reqs[i].pulse_key, []).append(i)
        for key, members in by_key.items():
            if len(members) > 1 or not reqs[members[0]].job.tables:
                te = max(reqs[i].job.t_end for i in members)
                tabs = reqs[members[0]].table_maker(te)
                for i in members:
                    reqs[i].job.tables = tabs
                    # ... but every run sees only the samples of its OWN pulse file (reference :213), end value held
                    if reqs[i].sampled:
                        reqs[i].job.table_len = len(np.arange(reqs[i].job.t_start, reqs[i].job.t_end, step=reqs[i].job.dt))
        plain = [i for i in idx if not reqs[i].calc_dynmap]
        # pulse files (read_pulse_file) bring their own grid: one batch per table grid
        by_grid: Dict[tuple, List[int]] = {}
        for i in plain:
            g = next(((round(tb.t0, 9), round(tb.dt, 12)) for tb in reqs[i].job.tables.values() if tb is not None), None)
            by_grid.setdefault(g, []).append(i)
        for members in by_grid.values():
            from pyaceqd_b200 import distributed as _dist
            jobs = [reqs[i].job for i in members]
            if tail_reduce is not None:
                outs = eng.run_jobs(prob, pt, jobs, tail_reduce=tail_reduce)
                for i, o in zip(members, outs):
                    results[i] = o
                continue
            if distributed and _dist.is_multi_rank() and len(jobs) > 1:
                # the sweep shards over the ranks (one GPU each) and is all-gathered once; every rank returns the
                # full result like wait(futures) in the reference
                outs = _dist.run_jobs_sharded(eng, prob, pt, jobs)
            else:
                outs = eng.run_jobs(prob, pt, jobs)
            for i, o in zip(members, outs):
                results[i] = _finish(reqs[i], o)
        dyn = [i for i in idx if reqs[i].calc_dynmap]
        if dyn:
            for i, o in zip(dyn, _run_dynmaps(eng, [reqs[i] for i in dyn])):
                results[i] = o
    for r, res in zip(reqs, results):
        r.result = res
    return results


class SweepResult:
    """Results of :func:`run_sweep_arrays`: ``res[i]`` is what the ``i``-th ``system(...)`` call returns (time row +
    one row per output operator, only the job's kept rows); :meth:`last_rows` gives the final row of every job as
    one ``[n_jobs, n_out]`` array without building them."""

    def __init__(self, out, out_off, n_rows, n_out, t_start, dt, n_steps):
        self.out, self.out_off, self.n_rows, self.n_out = out, out_off, n_rows, n_out
        self.t_start, self.dt, self.n_steps = t_start, dt, n_steps

    def __len__(self):
        return len(self.out_off)

    def __getitem__(self, i):
        r, o = int(self.n_rows[i]), int(self.out_off[i])
        res = np.empty((1 + self.n_out, r), dtype=complex)
        res[0] = self.t_start + self.dt * np.arange(int(self.n_steps[i]) + 1 - r, int(self.n_steps[i]) + 1)
        res[1:] = self.out[o: o + r * self.n_out].reshape(r, self.n_out).T
        return res

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def last_rows(self):
        idx = (self.out_off + (self.n_rows - 1) * self.n_out)[:, None] + np.arange(self.n_out)[None, :]
        return self.out[idx]


def run_sweep_arrays(template: Request, t_end, mto_templates, mto_times, tails, tail_reduce=None, distributed=None):
    """A sweep whose calls differ from ``template`` (the deferred first call) only in ``t_end`` and in the times of
    the SAME multi-time operators (``mto_times[J, M]``, NaN = absent): planned from arrays
    (:func:`pyaceqd_b200.planner.arrays_from_sweep`), no per-call Python work.  Same results as submitting the calls
    one by one, including the per-run length of sampled pulse files (reference ``:213``)."""
    from pyaceqd_b200 import planner
    from pyaceqd_b200.engine import default_engine
    eng = default_engine()
    prob, pt, job0 = template.problem, template.pt, template.job
    t_end = np.asarray(t_end, dtype=float)
    tables = template.table_maker(float(t_end.max()))
    table_len = None
    if template.sampled:          # len(np.arange(t_start, t_end, dt)) of every run
        table_len = np.ceil((t_end - job0.t_start) / job0.dt).astype(np.int64)
    parsed = prob.parse_mtos([dict(m, time=0.0) for m in mto_templates])
    arr = planner.arrays_from_sweep(prob, dt=job0.dt, t_start=job0.t_start, t_end=t_end,
                                    superops=[m.superop for m in parsed], before=[m.before for m in parsed],
                                    mto_times=mto_times, tails=tails, tables=tables, table_len=table_len)
    from pyaceqd_b200 import distributed as _dist
    if distributed is None:
        distributed = os.environ.get("ACEQD_DISTRIBUTED", "0") == "1"
    if distributed and tail_reduce is None and _dist.is_multi_rank() and arr.n_jobs > 1:
        # the sweep shards over the ranks (one GPU each) and is all-gathered once; every rank returns the full result
        res = _dist.run_arrays_sharded(eng, prob, pt, arr)
    else:
        res = eng.run_arrays(prob, pt, arr, tail_reduce=tail_reduce)
    if tail_reduce is not None:
        return res
    out, out_off, n_rows = res
    return SweepResult(out, out_off, n_rows, prob.n_out, job0.t_start, job0.dt, arr.n_steps)


def _run_dynmaps(eng, reqs: List[Request]):
    """``DynamicalMap.E`` (reference ``:328-335``) for requests of one (problem, PT) group: the physical runs in one
    batch, the NL unit vectors of EVERY request as initial states with identity outputs in a second one.  Layout as
    the reference's consumers expect (``tools.py:470-479``): ``E[i] = E_{t_{i+1}, t_0}``, i.e. ``E[0] rho0 =
    rho(t_1)`` -- the identity at ``t_0`` is not part of the array unless ``constants.dynmap_includes_t0`` is set."""
    import copy
    prob, pt = reqs[0].problem, reqs[0].pt
    NL = prob.NL
    phys = [_finish(r, o) for r, o in zip(reqs, eng.run_jobs(prob, pt, [r.job for r in reqs]))]
    key = ("dynmap", id(prob))
    with _cache_lock:
        basis = _problem_cache.get(key)
        if basis is None:
            basis = copy.copy(prob)
            basis.out_w = np.eye(NL, dtype=complex)
            basis.meta = dict(prob.meta)
            _problem_cache[key] = basis
    jobs = []
    for r in reqs:
        for j in range(NL):
            jb = copy.copy(r.job)
            jb.rho0 = np.zeros(NL, dtype=complex)
            jb.rho0[j] = 1.0
            jb.tail_rows = 0
            jobs.append(jb)
    cols = eng.run_jobs(basis, pt, jobs)
    out = []
    for k, r in enumerate(reqs):
        E = np.empty((r.job.n_steps + 1, NL, NL), dtype=complex)
        for j in range(NL):
            E[:, :, j] = cols[k * NL + j].T
        out.append((phys[k], (E if getattr(constants, "dynmap_includes_t0", False) else E[1:])))
    return out
