"""Shared machinery of the batch issuers.

Every (t, tau) workflow of the reference is one of two sweep shapes, written out many times
with a ``ThreadPoolExecutor`` of ACE subprocesses:

* **tail sweep** -- one job per ``t1``: multi-time operators moved to ``t1`` (+ offsets), the
  last rows of one output give the ``tau > 0`` axis and one row of a second output the
  ``tau = 0`` element (``two_time/correlations.py:153-184``, ``two_time/G1.py:66-89``,
  ``pol_entanglement/G2.py:189-204,488-530``, ``timebin/twophoton_new.py:215-262``);
* **triangular sweep** -- one job per pair ``t1 <= t2`` with three operators and only the very
  last output value kept (``timebin/twophoton_new.py:515-557``).

Here each sweep is ONE deferred GPU batch (:class:`~pyaceqd_b200.batch.BatchExecutor`): jobs that
share the drive are forked from a common trunk and only the rows a consumer reads are copied back.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from pyaceqd_b200.batch import BatchExecutor, wait
import pyaceqd_b200.constants as constants


def at_time(mto: dict, time: float) -> dict:
    """Copy of a multitime dict with its ``time`` set (callers must not share dicts between jobs)."""
    m = dict(mto)
    m["time"] = time
    return m


def run_sweep(system, jobs: Sequence[dict], *pulses, options: Optional[dict] = None, workers=None, tail_reduce=None) -> List:
    """Submit ``system(t0, tend, *pulses, multitime_op=..., output_ops=..., suffix=i, **options)`` for
    every job spec ``{"t0", "tend", "mtos", "output_ops", "tail"}`` in one batch; return results.
    ``tail_reduce = (pairs, spacing)``: per job the tau integrals of its tail (reduced on the device) instead."""
    opts = dict(options or {})
    uniform = _uniform_sweep(jobs)
    if uniform is not None:      # the usual case: the jobs differ in numbers only -> rows of arrays, no per-job call
        t0, tends, templates, times, output_ops, tails = uniform
        res = run_sweep_arrays(system, t0, tends, templates, times, *pulses, output_ops=output_ops, tails=tails,
                               options=opts, workers=workers, tail_reduce=tail_reduce, _generic=False)
        if res is not None:
            return list(res)
    with BatchExecutor(max_workers=workers, tail_reduce=tail_reduce) as ex:
        futs = []
        for i, jb in enumerate(jobs):
            kw = dict(opts)
            if jb.get("output_ops") is not None:
                kw["output_ops"] = jb["output_ops"]
            futs.append(ex.submit_tail(jb.get("tail", 0), system, jb.get("t0", 0), jb["tend"], *pulses,
                                       multitime_op=jb["mtos"], suffix=i, **kw))
        wait(futs)
    return [f.result() for f in futs]


def _uniform_sweep(jobs):
    """``(t0, tends, operator templates, times[J, M], output_ops, tails)`` if all jobs apply the same operators (in
    the same file order) to the same outputs from the same start, else None."""
    if len(jobs) < 2:
        return None
    as_list = lambda m: [m] if isinstance(m, dict) else list(m or [])
    first = as_list(jobs[0]["mtos"])
    sig = [(m.get("operator"), m.get("applyFrom", ""), str(m.get("applyBefore", "false")).strip().lower()) for m in first]
    t0, outs = jobs[0].get("t0", 0), jobs[0].get("output_ops")
    times = np.empty((len(jobs), len(first)))
    for i, jb in enumerate(jobs):
        ms = as_list(jb["mtos"])
        if len(ms) != len(sig) or jb.get("t0", 0) != t0 or jb.get("output_ops") != outs:
            return None
        for k, m in enumerate(ms):
            if (m.get("operator"), m.get("applyFrom", ""), str(m.get("applyBefore", "false")).strip().lower()) != sig[k] \
                    or "time" not in m:
                return None
            times[i, k] = m["time"]
    templates = [{k: v for k, v in m.items() if k != "time"} for m in first]
    return (t0, np.asarray([jb["tend"] for jb in jobs], dtype=float), templates, times, outs,
            np.asarray([jb.get("tail", 0) or 0 for jb in jobs], dtype=np.int64))


def run_sweep_arrays(system, t0, tends, mto_templates, mto_times, *pulses, output_ops=None, tails=0,
                     options: Optional[dict] = None, workers=None, tail_reduce=None, _generic=True):
    """The same sweep as :func:`run_sweep` for jobs that differ only in numbers: job ``i`` is
    ``system(t0, tends[i], *pulses, multitime_op=[dict(m, time=mto_times[i][k]) for k, m in enumerate(mto_templates)],
    output_ops=output_ops, **options)`` (a NaN time drops the operator).  Only the first call goes through the
    adapter; the others are rows of arrays (``general_system.run_sweep_arrays``).  An adapter that does not defer
    (it post-processes its result) gets the calls one by one, like :func:`run_sweep`."""
    from pyaceqd_b200.general_system import general_system as gs
    tends = np.asarray(tends, dtype=float)
    mto_times = np.asarray(mto_times, dtype=float).reshape(len(tends), len(mto_templates))
    tails_arr = np.broadcast_to(np.asarray(tails, dtype=np.int64), tends.shape)
    opts = dict(options or {})
    if output_ops is not None:
        opts["output_ops"] = output_ops

    def mtos_of(i):
        return [at_time(m, float(t)) for m, t in zip(mto_templates, mto_times[i]) if not np.isnan(t)]

    sink: list = []
    prev = getattr(gs._capture, "sink", None)
    gs._capture.sink = sink
    try:
        first = system(t0, float(tends[0]), *pulses, multitime_op=mtos_of(0), suffix=0, **opts)
    except Exception:      # noqa: BLE001 - an adapter that chokes on the placeholder: the generic route decides
        first = None
    finally:
        gs._capture.sink = prev
    if not isinstance(first, gs.Request) or first.calc_dynmap:
        if not _generic:
            return None
        jobs = [{"t0": t0, "tend": float(tends[i]), "mtos": mtos_of(i), "tail": int(tails_arr[i])} for i in range(len(tends))]
        return run_sweep(system, jobs, *pulses, options=opts, workers=workers, tail_reduce=tail_reduce)
    return gs.run_sweep_arrays(first, tends, mto_templates, mto_times, tails_arr, tail_reduce=tail_reduce)


def tail_series(res, n_after: int, i_tau: int = 1, i_zero: int = 2) -> np.ndarray:
    """``[G(tau=0), G(tau_1), ..., G(tau_n)]`` from one job result: the ``n_after`` last rows of output
    ``i_tau`` and, for ``tau = 0``, the row just before them of output ``i_zero`` -- the full
    operator product evaluated at the MTO time."""
    col = np.empty(n_after + 1, dtype=complex)
    col[0] = res[i_zero][-(n_after + 1)]
    if n_after > 0:
        col[1:] = res[i_tau][-n_after:]
    return col


def symmetrised_spectrum(t_axis, tau_axis, g1, hbar: float = constants.hbar):
    """Emission spectrum from ``G1(t, tau)``: extend to negative ``tau`` by conjugation, FFT along
    ``tau`` for every ``t``, integrate over ``t`` (``two_time/G1.py:101-110``,
    ``pol_entanglement/G2.py:225-241``).  Returns ``(energies, spectrum, spectra[t, E])``."""
    n = len(tau_axis)
    dtau = abs(tau_axis[1] - tau_axis[0])
    energies = np.fft.fftshift(-2 * np.pi * hbar * np.fft.fftfreq(2 * n - 1, d=dtau))
    sym = np.concatenate([g1[:, ::-1], np.conj(g1[:, 1:])], axis=1)
    spectra = np.fft.fftshift(np.fft.fft(sym, axis=1), axes=1)
    spectrum = np.real(np.trapezoid(spectra, t_axis, axis=0))
    return energies, spectrum, spectra
