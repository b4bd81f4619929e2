"""Two-time correlation functions G(t, tau) for any ``system`` adapter.

Same entry points, arguments and return layout as the reference's
``pyaceqd/two_time/correlations.py:24-320`` (``two_op_one_time``, ``three_op_one_time``,
``two_op_two_time``, ``three_op_two_time``, ``five_op_two_time``).  The reference submits one ACE
subprocess per ``t`` to a ``ThreadPoolExecutor`` (``:153-170``); here the same ``submit`` calls go
to a :class:`~pyaceqd_b200.batch.BatchExecutor`, which turns the sweep into one GPU batch with
the common prefix propagated once (trunk/branch forking).

Result conventions kept verbatim: the ``tau = 0`` column is read from the
second output operator (the full operator product) at the MTO time, the ``tau > 0`` columns from
the first output operator over the last ``n_tau`` rows.
"""
from __future__ import annotations

import numpy as np

from pyaceqd_b200.batch import BatchExecutor, wait


def _product(*ops):
    return "(" + "*".join(ops) + ")"


def _ops_one_time(system, *pulses, t0=-500, t_MTO=0, tend=500, dt=0.1,
                  options={"lindblad": True, "phonons": False}, debug=False):
    """One trajectory with MTOs already in ``options``; reference ``:24-52``."""
    t, out_b, out_0 = system(t0, tend, *pulses, dt=dt, **options)
    t = np.round(t, 6)
    n_tau = int((tend - t_MTO) / dt) + 1
    tau = np.linspace(t_MTO, tend, n_tau)
    i_mto = np.where(t == t_MTO)[0][0]
    g = np.empty(n_tau, dtype=complex)
    g[0] = out_0[i_mto]
    g[1:] = out_b[i_mto + 1:]
    return tau, g


def two_op_one_time(system, *pulses, opA="|1><0|_2", opB="|0><1|_2", t0=-500, t_MTO=0, tend=500, dt=0.1,
                    options={"lindblad": True, "phonons": False}, debug=False):
    """``<A(t_MTO + tau) B(t_MTO)>`` (e.g. G1(tau)); reference ``:54-91``."""
    options["output_ops"] = [opA, _product(opA, opB)]
    options["multitime_op"] = [{"operator": opB, "applyFrom": "_left", "applyBefore": "false", "time": t_MTO}]
    return _ops_one_time(system, *pulses, t0=t0, t_MTO=t_MTO, tend=tend, dt=dt, options=options, debug=debug)


def three_op_one_time(system, *pulses, opA="|1><0|_2", opB="|1><1|_2", opC="|0><1|_2", t0=-500, t_MTO=0, tend=500,
                      dt=0.1, options={"lindblad": True, "phonons": False}, debug=False):
    """``<A(t_MTO) B(t_MTO + tau) C(t_MTO)>`` (e.g. G2(tau)); reference ``:93-133``."""
    options["output_ops"] = [opB, _product(opA, opB, opC)]
    options["multitime_op"] = [
        {"operator": opA, "applyFrom": "_right", "applyBefore": "false", "time": t_MTO},
        {"operator": opC, "applyFrom": "_left", "applyBefore": "false", "time": t_MTO}]
    return _ops_one_time(system, *pulses, t0=t0, t_MTO=t_MTO, tend=tend, dt=dt, options=options, debug=debug)


def _ops_two_time(system, t_axis, *pulses, mtos=[], tau_max=500, dt=0.1,
                  options={"lindblad": True, "phonons": False}, debug=False, workers=15, n_mto=None, t_start=0):
    """One trajectory per ``t1`` in ``t_axis`` with the first ``n_mto`` MTOs moved to ``t1``;
    reference ``:135-184``.  Returns ``t1, tau, G[len(t1), n_tau + 1]``."""
    if len(mtos) < n_mto:
        raise ValueError("multi-time operators are required for the two-time correlation function.")
    if t_start > 0:
        raise ValueError("t_start > 0 is not supported yet. Use t_start<=0 to e.g. reach a stationary state "
                         "before applying the MTO.")
    fixed = [dict(m) for m in mtos[n_mto:]]
    t1 = t_axis
    n_tau = int(tau_max / dt)
    tau = np.linspace(0, tau_max, n_tau + 1)
    G = np.empty((len(t1), len(tau)), dtype=complex)
    with BatchExecutor(max_workers=workers) as executor:
        futures = []
        for i, t1_i in enumerate(t1):
            moving = []
            for m in mtos[:n_mto]:
                m = dict(m)
                m["time"] = t1_i
                moving.append(m)
            futures.append(executor.submit(system, t_start, t1_i + tau_max, *pulses, dt=dt, suffix=i,
                                           multitime_op=moving + [dict(m) for m in fixed], **options))
        wait(futures)
    for j, f in enumerate(futures):
        res = f.result()
        G[j, 1:] = res[1][-n_tau:]
        G[j, 0] = res[2][-(n_tau + 1)]
    return t1, tau, G


def two_op_two_time(system, t_axis, *pulses, opA="|1><0|_2", opB="|0><1|_2", tau_max=500, dt=0.1,
                    options={"lindblad": True, "phonons": False}, debug=False, workers=15):
    """``<A(t + tau) B(t)>`` (e.g. G1(t, tau)); reference ``:186-225``."""
    options["output_ops"] = [opA, _product(opA, opB)]
    mtos = [{"operator": opB, "applyFrom": "_left", "applyBefore": "false"}]
    return _ops_two_time(system, t_axis, *pulses, mtos=mtos, tau_max=tau_max, dt=dt, options=options,
                         debug=debug, workers=workers, n_mto=1)


def three_op_two_time(system, t_axis, *pulses, opA="|1><0|_2", opB="|1><1|_2", opC="|0><1|_2", tau_max=500,
                      dt=0.1, t_start=0, options={"lindblad": True, "phonons": False}, debug=False, workers=15):
    """``<A(t) B(t + tau) C(t)>`` (e.g. G2(t, t + tau)); reference ``:227-270``."""
    options["output_ops"] = [opB, _product(opA, opB, opC)]
    mtos = [{"operator": opA, "applyFrom": "_right", "applyBefore": "false"},
            {"operator": opC, "applyFrom": "_left", "applyBefore": "false"}]
    return _ops_two_time(system, t_axis, *pulses, mtos=mtos, tau_max=tau_max, dt=dt, options=options,
                         debug=debug, workers=workers, n_mto=2, t_start=t_start)


def five_op_two_time(system, t_axis, *pulses, opA="|1><0|_2", opB="|1><0|_2", opC="|1><1|_2", opD="|0><1|_2",
                     opE="|0><1|_2", tau_max=500, dt=0.1, t_start=-500,
                     options={"lindblad": True, "phonons": False}, debug=False, workers=15):
    """``<A(0) B(t) C(t + tau) D(t) E(0)>``; reference ``:272-320`` (including its documented
    caveat: the (t=0, tau=0) element uses ``<B C D>`` only)."""
    options["output_ops"] = [opC, _product(opA, opB, opC, opD, opE)]
    mtos = [{"operator": opB, "applyFrom": "_right", "applyBefore": "false"},
            {"operator": opD, "applyFrom": "_left", "applyBefore": "false"},
            {"operator": opA, "applyFrom": "_right", "applyBefore": "false", "time": 0},
            {"operator": opE, "applyFrom": "_left", "applyBefore": "false", "time": 0}]
    return _ops_two_time(system, t_axis, *pulses, mtos=mtos, tau_max=tau_max, dt=dt, options=options,
                         debug=debug, workers=workers, n_mto=2, t_start=t_start)


def G2_spectral_integral(t1, tau, G):
    """Time-integrated second-order correlation ``int dt int dtau G(t, tau)`` on the given axes
    (what the consumers of ``three_op_two_time`` compute, e.g. ``pol_entanglement/G2.py:292-299``)."""
    return np.trapezoid(np.trapezoid(G, tau, axis=1), t1)
