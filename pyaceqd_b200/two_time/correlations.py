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
from pyaceqd_b200.sweeps import run_sweep_arrays


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
    t1 = np.asarray(t_axis, dtype=float)
    n_tau = int(tau_max / dt)
    tau = np.linspace(0, tau_max, n_tau + 1)
    # one call per t1, as the reference submits them -- but only the first goes through the adapter, the others are
    # rows of (end time, operator times): file order of a call's operator list = moving ones first, then the fixed
    templates = [dict(m) for m in mtos[:n_mto]] + [dict(m) for m in mtos[n_mto:]]
    times = np.empty((len(t1), len(templates)))
    times[:, :n_mto] = t1[:, None]
    for k, m in enumerate(mtos[n_mto:]):
        times[:, n_mto + k] = m["time"]
    res = run_sweep_arrays(system, t_start, t1 + tau_max, templates, times, *pulses, tails=n_tau + 1,
                           options=dict(options, dt=dt), workers=workers)
    G = np.empty((len(t1), len(tau)), dtype=complex)
    for j, r in enumerate(res):
        G[j, 1:] = r[1][-n_tau:]
        G[j, 0] = r[2][-(n_tau + 1)]
    return t_axis, tau, G


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


def get_spectrum(g1, tau, dir="", plot=False):
    """Spectrum under continuous-wave excitation from ``G1(tau)`` (reference ``:322-382``): the stationary offset
    ``g1[-1]`` is removed, negative delays are the conjugates, then an FFT.  Returns ``(s_omega, energies)`` in meV,
    both fft-shifted."""
    import pyaceqd_b200.constants as constants
    g1 = np.array(g1, dtype=complex)
    dtau = np.abs(tau[1] - tau[0])
    g1 = g1 - g1[-1]
    g1 = np.concatenate((np.conj(np.flip(g1[1:])), g1))
    tau = np.concatenate((-np.flip(tau[1:]), tau))
    s_omega = np.fft.fftshift(np.real(np.fft.fft(g1)))
    energies = np.fft.fftshift(2 * np.pi * constants.hbar * np.fft.fftfreq(len(g1), d=dtau))
    if plot:
        import matplotlib.pyplot as plt
        for name, x, y, lim, labels in (
                ("g1_tendsymm.png", tau, np.abs(g1), (-1, 1), ("Time (ps)", "|G1(t)|")),
                ("spectrum_log.png", energies, np.log(np.abs(s_omega)), (-3, 3), ("Frequency (meV)", "S(omega)")),
                ("spectrum_nolog.png", energies, np.abs(s_omega), (-3, 3), ("Frequency (meV)", "S(omega)"))):
            plt.clf()
            plt.plot(x, y)
            plt.xlim(*lim)
            plt.xlabel(labels[0])
            plt.ylabel(labels[1])
            plt.savefig(dir + name)
        plt.clf()
    return s_omega, energies


def G2_spectral_integral(t1, tau, G):
    """Time-integrated second-order correlation ``int dt int dtau G(t, tau)`` on the given axes
    (what the consumers of ``three_op_two_time`` compute, e.g. ``pol_entanglement/G2.py:292-299``)."""
    return np.trapezoid(np.trapezoid(G, tau, axis=1), t1)


# ------------------------------------------------------------------------------------ time-local maps
def _tl_correlation(system, t_axis, pulses, left, right, out, tau0, t_mem, tau_max, dt, rho0, options, use_dm,
                    fortran_args):
    """Shared body of the ``tl_*_two_time`` functions (reference ``:450-616,696-864``).

    ``G[i, 0] = Tr(tau0 rho(t_i))`` and ``G[i, k] = Tr(out E_k[left rho(t_i) right])`` where ``E_k`` is the
    propagation over ``k`` steps after ``t_i``: with ``use_dm`` the time-local maps of the whole window
    (one dynamical-map run, then matrix-vector chains on the GPU chain kernel), otherwise -- valid
    for time-independent dynamics only -- powers of the stationary map extracted after ``t_mem``.
    ``fortran_args`` switches to the reference's ``fortran_only`` call of ``calc_onetime_parallel``
    with its column-major conventions (``:534,782``)."""
    from pyaceqd_b200.tlmap import Programs
    from pyaceqd_b200.tools import calc_tl_dynmap_pseudo, extract_dms
    from pyaceqd_b200.two_time import propagate_tau_module
    if not t_axis[0] == 0:
        raise ValueError("t_axis must start at 0.")
    dim = len(rho0[0])
    NL = dim * dim
    n_tau = int(tau_max / dt)
    tau = np.linspace(0, tau_max, n_tau + 1)
    v0 = np.asarray(rho0, dtype=complex).reshape(NL)
    start = np.kron(left, right.T)                     # row-major: vec(L rho R) = (L (x) R^T) vec(rho)
    w_out, w_tau0 = out.T.reshape(-1), tau0.T.reshape(-1)
    if use_dm:
        result, dm = system(0, t_axis[-1] + tau_max, *pulses, dt=dt, rho0=rho0, multitime_op=[], calc_dynmap=True,
                            **options)
        t_sim = np.round(result[0].real, 6)
        tl = calc_tl_dynmap_pseudo(dm, t_sim)
        if fortran_args is not None:
            return t_axis, tau, propagate_tau_module.calc_onetime_parallel(
                np.asfortranarray(tl.transpose(1, 2, 0)), v0, n_tau, dim, *fortran_args, t_sim, t_axis)
        pr = Programs(NL)
        off = pr.add(tl)
        G = np.zeros((len(t_axis), n_tau + 1), dtype=complex)
        v, j = v0.copy(), 0
        for i, t in enumerate(t_axis):
            while t_sim[j] < t:
                v = tl[j] @ v
                j += 1
            G[i, 0] = w_tau0 @ v
            pr.chain(start @ v, [(off + j, n_tau, 1, 1)])
        out_vals, _ = pr.run(w=w_out[None])
        G[:, 1:] = out_vals[:, :n_tau, 0]
        return t_axis, tau, G
    if options.get("phonons"):
        print("phonons not implemented yet")
        return t_axis, tau, np.zeros((len(t_axis), n_tau + 1), dtype=complex)
    # stationary map from a short dynamical-map run (the operators at 2 t_mem are part of the reference's
    # call but do not enter the stationary map, which is read off before them)
    result, dm = system(0, 4 * t_mem, *pulses, dt=dt, rho0=rho0, multitime_op=[], calc_dynmap=True, **options)
    t_sim = np.round(result[0].real, 6)
    tl_map, _ = extract_dms(calc_tl_dynmap_pseudo(dm, t_sim), t_sim, t_mem, [2 * t_mem])
    pr = Programs(NL)
    off = pr.add(tl_map)
    G = np.zeros((len(t_axis), n_tau + 1), dtype=complex)
    v = v0.copy()
    for i, t in enumerate(t_axis):
        if i > 0:
            v = np.linalg.matrix_power(tl_map, int((t - t_axis[i - 1]) / dt)) @ v
        G[i, 0] = w_tau0 @ v
        pr.chain(start @ v, [(off, n_tau, 1, 0)])     # column k = after k applications of the stationary map
    out_vals, _ = pr.run(w=w_out[None])
    # (the reference's stationary branch shifts the tau axis by one step and repeats the first `dim`
    #  entries, an artefact of passing a matrix to tl_pad_stationary_nsteps, :612-614; the layout here
    #  is the one of its use_dm branch, :583-587)
    G[:, 1:] = out_vals[:, :n_tau, 0]
    return t_axis, tau, G


def tl_two_op_two_time(system, t_axis, *pulses, t_mem=10, opA="|1><0|_2", opB="|0><1|_2", tau_max=500, dt=0.1,
                       rho0=np.array([[1, 0], [0, 0]], dtype=complex), options={"lindblad": True, "phonons": False},
                       debug=False, workers=15, use_dm=False, fortran_only=False):
    """``<A(t + tau) B(t)>`` from time-local dynamical maps (reference ``:450-616``)."""
    from pyaceqd_b200.tools import op_to_matrix
    A, B = op_to_matrix(opA), op_to_matrix(opB)
    I = np.identity(len(rho0[0]))
    return _tl_correlation(system, t_axis, pulses, B, I, A, A @ B, t_mem, tau_max, dt, rho0, options, use_dm,
                           (I, A, B) if fortran_only else None)


def tl_three_op_two_time(system, t_axis, *pulses, t_mem=10, opA="|1><0|_2", opB="|1><1|_2", opC="|0><1|_2",
                         tau_max=500, dt=0.1, rho0=np.array([[1, 0], [0, 0]], dtype=complex),
                         options={"lindblad": True, "phonons": False}, debug=False, workers=15, use_dm=False,
                         fortran_only=False):
    """``<A(t) B(t + tau) C(t)>`` from time-local dynamical maps (reference ``:696-864``)."""
    from pyaceqd_b200.tools import op_to_matrix
    A, B, C = op_to_matrix(opA), op_to_matrix(opB), op_to_matrix(opC)
    return _tl_correlation(system, t_axis, pulses, C, A, B, A @ B @ C, t_mem, tau_max, dt, rho0, options, use_dm,
                           (A, B, C) if fortran_only else None)


# ------------------------------------------------------------------------------------ time-local maps with phonons
def _phonon_maps(system, pulses, opA, opC, t_mem, dt, rho0, options):
    """The stationary pieces both phonon variants start from (reference ``:875-885,1022-1033``): one dynamical-map
    run over ``4 t_mem`` with the two operators applied at ``1.2 t_mem`` -- outside the memory of the start --
    gives the stationary map before the operators (``tl_map``), the explicit maps of the first memory time and of
    the memory time after the operators (``blocks``; the operators are part of ``blocks[1][0]``), and the
    stationary map after them (``tl_map2``, the last step of the run)."""
    from pyaceqd_b200.tools import calc_tl_dynmap_pseudo, extract_dms
    mtos = [{"operator": opC, "applyFrom": "_left", "applyBefore": "false", "time": 1.2 * t_mem},
            {"operator": opA, "applyFrom": "_right", "applyBefore": "false", "time": 1.2 * t_mem}]
    result, dm = system(0, 4 * t_mem, *pulses, dt=dt, rho0=rho0, multitime_op=mtos, calc_dynmap=True, **options)
    t_sim = np.round(result[0].real, 6)
    tl = calc_tl_dynmap_pseudo(dm, t_sim)
    tl_map, blocks = extract_dms(tl, t_sim, t_mem, [np.round(1.2 * t_mem, 6)])
    return tl_map, np.array(blocks, dtype=complex), tl[-1]


def _moved(opA, opC, t):
    return [{"operator": opC, "applyFrom": "_left", "applyBefore": "false", "time": t},
            {"operator": opA, "applyFrom": "_right", "applyBefore": "false", "time": t}]


def tl_three_op_two_time_phonons(system, t_axis, *pulses, t_mem=10, opA="|1><0|_2", opB="|1><1|_2", opC="|0><1|_2",
                                 tau_max=500, dt=0.1, rho0=np.array([[1, 0], [0, 0]], dtype=complex),
                                 options={"lindblad": True, "phonons": True}, debug=False, fortran_only=False):
    """``<A(t) B(t + tau) C(t)>`` with a phonon memory of ``t_mem`` from time-local maps (reference ``:866-1011``).

    ``rho(t)`` comes from the explicit maps of the first memory time, then powers of the stationary map; the
    ``tau`` axis from the explicit maps after the operators -- for ``t < t_mem`` those of a run with the operators
    AT ``t`` (the start still inside the memory), else the stationary block -- then powers of ``tl_map2``.
    All runs below ``t_mem`` are ONE batch, all ``tau`` chains one launch of the chain kernel."""
    from pyaceqd_b200.tlmap import Programs
    from pyaceqd_b200.tools import calc_tl_dynmap_pseudo, extract_dms, op_to_matrix
    if not t_axis[0] == 0:
        raise ValueError("t_axis must start at 0.")
    t_axis = np.round(t_axis, 6)
    A, B, C = op_to_matrix(opA), op_to_matrix(opB), op_to_matrix(opC)
    tl_map, blocks, tl_map2 = _phonon_maps(system, pulses, opA, opC, t_mem, dt, rho0, options)
    n_tauc = blocks.shape[1]
    n_tau = int(tau_max / dt)
    tau = np.linspace(0, tau_max, n_tau + 1)
    dim = len(rho0[0])
    NL = dim * dim
    inside = np.where(t_axis < t_mem)[0]
    with BatchExecutor() as ex:
        futs = [ex.submit(system, 0, t_axis[i] + t_mem + 10 * dt, *pulses, dt=dt, rho0=rho0,
                          multitime_op=_moved(opA, opC, t_axis[i]), calc_dynmap=True, suffix=int(i), **options)
                for i in inside]
        wait(futs)
    pr = Programs(NL)
    off_stat, off_tl2 = pr.add(blocks[1]), pr.add(tl_map2)
    off_own = []
    for i, f in zip(inside, futs):
        result, dm = f.result()
        t_sim = np.round(result[0].real, 6)
        _, own = extract_dms(calc_tl_dynmap_pseudo(dm, t_sim), t_sim, t_mem, [t_axis[i]])
        off_own.append(pr.add(own[1][:n_tauc]))
    # rho on the simulation grid: n_tauc - 1 explicit maps, then the stationary one (reference :960-972)
    k_of = [int(np.round(t / dt, 6)) for t in t_axis]
    states = np.empty((max(k_of) + 1, NL), dtype=complex)
    states[0] = np.asarray(rho0, dtype=complex).reshape(NL)
    for k in range(max(k_of)):
        states[k + 1] = (blocks[0][k] if k < n_tauc - 1 else tl_map) @ states[k]
    G = np.zeros((len(t_axis), n_tau + 1), dtype=complex)
    w_abc, w_b = (A @ B @ C).T.reshape(-1), B.T.reshape(-1)
    for i, k in enumerate(k_of):
        G[i, 0] = w_abc @ states[k]
        first = off_own[i] if i < len(inside) else off_stat
        n1 = min(n_tauc, n_tau)
        pr.chain(states[k], [(first, n1, 1, 1), (off_tl2, n_tau - n1, 1, 0)])
    out_vals, _ = pr.run(w=w_b[None])
    G[:, 1:] = out_vals[:, :n_tau, 0]
    return t_axis, tau, G


def tl_threeoptwotime_phonons_dm(system, t_axis, *pulses, t_mem=10, opA="|1><0|_2", opB="|1><1|_2", opC="|0><1|_2",
                                 tau_max=500, dt=0.1, rho0=np.array([[1, 0], [0, 0]], dtype=complex),
                                 options={"lindblad": True, "phonons": True}, debug=False, fortran_only=False):
    """As :func:`tl_three_op_two_time_phonons`, but for ``t <= t_mem`` the first ``t_mem`` of the ``tau`` axis is
    the full dynamical map of a run with the operators at ``t`` applied to ``rho0`` (reference ``:1013-1185``) --
    i.e. the trajectory itself, which is what is run here (one job per ``t``, ONE batch, instead of the NL + 1 of
    a map) -- and, beyond ``t_mem``, ALL explicit maps of the first memory time before the stationary one."""
    from pyaceqd_b200.tlmap import Programs
    from pyaceqd_b200.tools import op_to_matrix
    if not t_axis[0] == 0:
        raise ValueError("t_axis must start at 0.")
    t_axis = np.round(t_axis, 6)
    A, B, C = op_to_matrix(opA), op_to_matrix(opB), op_to_matrix(opC)
    tl_map, blocks, tl_map2 = _phonon_maps(system, pulses, opA, opC, t_mem, dt, rho0, options)
    n_tauc = blocks.shape[1]
    n_tau = int(tau_max / dt)
    tau = np.linspace(0, tau_max, n_tau + 1)
    dim = len(rho0[0])
    NL = dim * dim
    inside = np.where(t_axis <= t_mem)[0]
    opts = dict(options)
    opts["output_ops"] = [_product(opA, opB, opC)] + ["|%d><%d|_%d" % (c, r, dim) for r in range(dim) for c in range(dim)]
    with BatchExecutor() as ex:
        futs = [ex.submit(system, 0, t_axis[i] + t_mem, *pulses, dt=dt, rho0=rho0,
                          multitime_op=_moved(opA, opC, t_axis[i]), suffix=int(i), **opts) for i in inside]
        wait(futs)
    pr = Programs(NL)
    off_stat, off_tl2 = pr.add(blocks[1]), pr.add(tl_map2)
    G = np.zeros((len(t_axis), n_tau + 1), dtype=complex)
    w_abc, w_b = (A @ B @ C).T.reshape(-1), B.T.reshape(-1)
    for i, f in zip(inside, futs):
        res = f.result()
        k = int(np.round(t_axis[i] / dt, 6))
        rho = res[2:2 + NL].T                      # rho[k] = vec(rho(t_k)), row-major (output |c><r| reads rho_rc)
        G[i, 0] = res[1][k]                        # the operator product at the operator time (before they act)
        n_map = min(rho.shape[0] - 1 - k, n_tau)
        G[i, 1:1 + n_map] = rho[k + 1:k + 1 + n_map] @ w_b
        pr.chain(rho[k + n_map], [(off_tl2, n_tau - n_map, 1, 0)])
    v = np.asarray(rho0, dtype=complex).reshape(NL)
    for m in blocks[0]:
        v = m @ v
    k_done = n_tauc
    for i in range(len(inside), len(t_axis)):
        k = int(np.round(t_axis[i] / dt, 6))
        v = np.linalg.matrix_power(tl_map, k - k_done) @ v
        k_done = k
        G[i, 0] = w_abc @ v
        n1 = min(n_tauc, n_tau)
        pr.chain(v, [(off_stat, n1, 1, 1), (off_tl2, n_tau - n1, 1, 0)])
    out_vals, _ = pr.run(w=w_b[None])
    for i in range(len(t_axis)):
        if i < len(inside):
            n_map = min(int(np.round((t_axis[i] + t_mem) / dt, 6)) - int(np.round(t_axis[i] / dt, 6)), n_tau)
            G[i, 1 + n_map:] = out_vals[i, :n_tau - n_map, 0]
        else:
            G[i, 1:] = out_vals[i, :n_tau, 0]
    return t_axis, tau, G
