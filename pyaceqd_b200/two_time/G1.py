"""First-order correlation ``G1(t, tau) = <sigma^+(t + tau) sigma(t)>`` and pulsed-Mollow spectra.

Entry points, arguments and return layout of the reference's ``pyaceqd/two_time/G1.py``
(``G1_twols`` ``:15-33``, ``G1_general`` ``:36-89``, ``pulsed_mollow_*`` ``:91-183``,
``simple_vhom`` ``:185-199``).  The reference submits one ACE subprocess per ``t`` (``:66-75``); here
the sweep is one GPU batch with the common prefix propagated once.
"""
from __future__ import annotations

import os

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.sweeps import at_time, run_sweep, symmetrised_spectrum
from pyaceqd_b200.tools import construct_t, export_csv
from pyaceqd_b200.two_level_system.tls import tls

HBAR = constants.hbar
temp_dir = constants.temp_dir


def G1_general(t0=0, tend=600, tau0=0, tauend=600, dt=0.1, dtau=0.02, *pulses, system=tls,
               multitime_op={"operator": "|0><1|_2", "applyFrom": "left"}, coarse_t=False, workers=10,
               prepare_only=False, simple_exp=False, gaussian_t=False, factor_tau=4, **options):
    """``t`` axis with step ``dt`` (coarsened away from the pulses if ``coarse_t``), ``tau`` axis with
    the simulation step ``dtau``; one trajectory per ``t`` ending at ``t + tauend``.  ``options`` must
    list two outputs: the ``tau = 0`` operator first, the ``tau > 0`` operator second (``:82-88``)."""
    t = np.linspace(t0, tend, int((tend - t0) / dt) + 1)
    n_tau = int((tauend - tau0) / dtau)
    tau = np.linspace(tau0, tauend, n_tau + 1)
    if coarse_t:
        # reference quirk kept: the first pulse binds to construct_t's positional dt_exp
        t = construct_t(t0, tend, dt, (3 if gaussian_t else 10) * dt, *pulses, factor_tau=factor_tau,
                        simple_exp=simple_exp, gaussian_t=bool(gaussian_t))
    if options.get("phonons") and prepare_only:
        system(0, 40, *pulses, dt=dtau, **options)     # builds and caches the process tensor
        return 0, 0, 0
    jobs = [{"t0": t0, "tend": ti + tauend, "mtos": at_time(multitime_op, ti), "tail": n_tau + 1} for ti in t]
    opts = dict(options)
    opts["dt"] = dtau
    res = run_sweep(system, jobs, *pulses, options=opts, workers=workers)
    g1 = np.zeros((len(t), len(tau)), dtype=complex)
    for i, r in enumerate(res):
        g1[i, 0] = r[1][-n_tau - 1]      # tau = 0 from the first output at the MTO time
        g1[i, 1:] = r[2][-n_tau:]        # tau > 0 from the second output
    return t, tau, g1


def G1_twols(t0=0, tend=600, tau0=0, tauend=600, dt=0.1, dtau=0.5, *pulses, ae=3.0, temperature=4, gamma_e=1 / 100,
             phonons=False, pt_file=None, workers=10, temp_dir=temp_dir, coarse_t=False, prepare_only=False,
             simple_exp=False, gaussian_t=False, factor_tau=4, **ops):
    """G1 of the two-level system: ``G1(t, 0) = <|1><1|>(t)``, ``G1(t, tau) = <|1><0|>(t + tau)`` after
    ``sigma = |0><1|`` acted from the left at ``t`` (reference ``:15-33``).  The drive is sampled once
    on the ``dtau`` grid into a pulse file shared by all trajectories, as in the reference."""
    grid = np.arange(t0, tend + tauend + dtau, step=dtau)
    field = np.zeros_like(grid, dtype=complex)
    for p in pulses:
        field = field + p.get_total(grid)
    pulse_file = temp_dir + "tls_G1_pulse_{}.dat".format(os.getpid())   # prefix concatenation, as the reference
    export_csv(pulse_file, grid, field.real, field.imag, precision=8, delimit=' ')
    options = {"gamma_e": gamma_e, "phonons": phonons, "ae": ae, "temperature": temperature, "lindblad": True,
               "pt_file": pt_file, "temp_dir": temp_dir, "pulse_file": pulse_file,
               "output_ops": ["|1><1|_2", "|1><0|_2"]}
    options.update(ops)
    mto = {"operator": "|0><1|_2", "applyFrom": "_left", "applyBefore": "false"}
    try:
        return G1_general(t0, tend, tau0, tauend, dt, dtau, *pulses, system=tls, multitime_op=mto,
                          coarse_t=coarse_t, workers=workers, prepare_only=prepare_only, simple_exp=simple_exp,
                          gaussian_t=gaussian_t, factor_tau=factor_tau, **options)
    finally:
        if os.path.exists(pulse_file):
            os.remove(pulse_file)


def _mollow_scan(values, make_pulse, label, tend, tauend, dt, dtau, save_dir, name, **g1_kwargs):
    """Time-integrated emission spectrum for every entry of ``values`` (shared body of the three
    ``pulsed_mollow_*`` functions, reference ``:91-183``)."""
    n_tau = int(tauend / dtau)
    spectra = np.zeros((len(values), 2 * n_tau + 1))
    energies = None
    for i, v in enumerate(values):
        t_axis, tau_axis, g1 = G1_twols(0, tend, 0, tauend, dt, dtau, make_pulse(v), coarse_t=True, **g1_kwargs)
        energies, spectra[i], _ = symmetrised_spectrum(t_axis, tau_axis, g1, HBAR)
        if save_dir is not None:     # progress survives an interrupted scan
            np.save(save_dir + "x" + name, energies)
            np.save(save_dir + "y" + name, np.asarray(values))
            np.save(save_dir + "z" + name, spectra)
    return energies, values, spectra


def pulsed_mollow_tls_pulses(pulse, areas, tend=500, tauend=500, dt=0.2, dtau=0.02, gamma_e=1 / 100, ae=3.0,
                             temperature=4, phonons=False, pt_file="tls_3.0nm_4k_th10_tmem20.48_dt0.02.ptr",
                             workers=7, temp_dir=temp_dir, save_dir=None, prepare_only=False, simple_exp=False,
                             gaussian_t=False, factor_tau=4):
    """Spectra of a given pulse object whose area ``e0`` is scanned over ``areas`` (reference ``:91-119``)."""
    def with_area(a):
        pulse.e0 = a
        return pulse
    name = "_tau{:.2f}_lifet{:.1f}_det{:.1f}.npy".format(pulse.tau, 1 / gamma_e, pulse.e_start)
    return _mollow_scan(areas, with_area, "area", tend, tauend, dt, dtau, save_dir, name, ae=ae, gamma_e=gamma_e,
                        phonons=phonons, workers=workers, temperature=temperature, pt_file=pt_file,
                        temp_dir=temp_dir, prepare_only=prepare_only, simple_exp=simple_exp,
                        gaussian_t=gaussian_t, factor_tau=factor_tau)


def pulsed_mollow_tls(pulse_tau, areas, detuning=0, tend=500, tauend=500, dt=0.2, dtau=0.02, gamma_e=1 / 100,
                      ae=3.0, temperature=4, phonons=False, pt_file="tls_3.0nm_4k_th10_tmem20.48_dt0.02.ptr",
                      workers=7, temp_dir=temp_dir, save_dir=None, prepare_only=False, simple_exp=False,
                      gaussian_t=False, **ops):
    """Pulsed Mollow spectra versus pulse area for Gaussian pulses of width ``pulse_tau`` (``:121-160``)."""
    name = "_tau{:.2f}_lifet{:.1f}_det{:.1f}.npy".format(pulse_tau, 1 / gamma_e, detuning)
    return _mollow_scan(areas, lambda a: ChirpedPulse(tau_0=pulse_tau, e_start=detuning, alpha=0, e0=a,
                                                      t0=pulse_tau * 4),
                        "area", tend, tauend, dt, dtau, save_dir, name, ae=ae, gamma_e=gamma_e, phonons=phonons,
                        workers=workers, temperature=temperature, pt_file=pt_file, temp_dir=temp_dir,
                        prepare_only=prepare_only, simple_exp=simple_exp, gaussian_t=gaussian_t, **ops)


def pulsed_mollow_energy(pulse_tau, detunings, area=3, tend=500, tauend=500, dt=0.2, dtau=0.02, gamma_e=1 / 100,
                         ae=3.0, temperature=4, phonons=False, pt_file="tls_3.0nm_4k_th10_tmem20.48_dt0.02.ptr",
                         workers=7, temp_dir=temp_dir, save_dir=None, prepare_only=False, simple_exp=False,
                         gaussian_t=False):
    """Pulsed Mollow spectra versus laser detuning at fixed area (reference ``:162-183``)."""
    name = "_tau{:.2f}_lifet{:.1f}_area{:.1f}.npy".format(pulse_tau, 1 / gamma_e, area)
    return _mollow_scan(detunings, lambda d: ChirpedPulse(tau_0=pulse_tau, e_start=d, alpha=0, e0=area,
                                                          t0=pulse_tau * 4),
                        "detuning", tend, tauend, dt, dtau, save_dir, name, ae=ae, gamma_e=gamma_e,
                        phonons=phonons, workers=workers, temperature=temperature, pt_file=pt_file,
                        temp_dir=temp_dir, prepare_only=prepare_only, simple_exp=simple_exp, gaussian_t=gaussian_t)


def simple_vhom(tend=600, tauend=600, dt=0.1, dtau=0.02, *pulses, ae=3.0, temperature=4, gamma_e=1 / 100,
                phonons=False, pt_file=None, workers=10, temp_dir=temp_dir, coarse_t=False, prepare_only=False):
    """Hong-Ou-Mandel visibility estimate ``2 int int |G1|^2 / brightness`` (reference ``:185-199``,
    marked "not tested" there)."""
    t, x = tls(0, tend, *pulses, dt=dtau, gamma_e=gamma_e, phonons=phonons, ae=ae, temperature=temperature,
               lindblad=True, pt_file=pt_file, temp_dir=temp_dir, output_ops=["|1><1|_2"])
    brightness = np.trapezoid(x, t)
    t, tau, g1 = G1_twols(0, tend, 0, tauend, dt, dtau, *pulses, ae=ae, temperature=temperature, gamma_e=gamma_e,
                          phonons=phonons, pt_file=pt_file, workers=workers, temp_dir=temp_dir, coarse_t=coarse_t,
                          prepare_only=prepare_only)
    return 2 * np.trapezoid(np.trapezoid(np.abs(g1) ** 2, t, axis=0), tau) / brightness
