"""Three-level bright/dark exciton model |0>=G, |1>=X, |2>=D (reference
``pyaceqd/two_level_system/reduced_dark.py:13-183``): the emitter alone, the emitter in a cavity, and the
early/late-bin integrals and ``G1`` sweeps built on the former -- each sweep one GPU batch."""
import os

import numpy as np

from pyaceqd_b200.general_system.general_system import system_ace_stream
import pyaceqd_b200.constants as constants
from pyaceqd_b200.tools import construct_t, export_csv, simple_t_gaussian

hbar = constants.hbar
temp_dir = constants.temp_dir


def darkmodel(t_start, t_end, *pulses, dt=0.5, delta_xd=0, gamma_e=1/65, phonons=False, ae=3.0, temperature=4,
              verbose=False, lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None,
              pulse_file_x=None, pulse_file_y=None, prepare_only=False,
              output_ops=["|0><0|_3", "|1><1|_3", "|2><2|_3"], initial="|0><0|_3"):
    """'x' polarisation couples G-D and D-X, 'y' couples G-X; only the bright exciton decays."""
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=20.48, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="tls_dark", threshold="10", threshold_ratio="0.3", buffer_blocksize="-1", dict_zero="16",
        precision="12", boson_e_max=7, system_op=["{}*|2><2|_3".format(-delta_xd)], pulse_file_x=pulse_file_x,
        pulse_file_y=pulse_file_y, boson_op="|1><1|_3 + |2><2|_3", initial=initial,
        lindblad_ops=[["|0><1|_3", gamma_e]] if lindblad else [],
        interaction_ops=[["|2><0|_3", "x"], ["|1><2|_3", "x"], ["|1><0|_3", "y"]], output_ops=output_ops,
        prepare_only=prepare_only)


def darkmodel_photons(t_start, t_end, *pulses, dt=0.1, delta_xd=0, delta_cx=-2, rad_loss=1/100, cav_loss=1/20,
                      cav_coupl=1/30, phonons=False, ae=3.0, temperature=4, verbose=False, lindblad=False,
                      temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None, pulse_file_x=None,
                      pulse_file_y=None, prepare_only=False,
                      output_ops=["|0><0|_3 otimes |0><0|_3", "|1><1|_3 otimes |0><0|_3", "|2><2|_3 otimes |0><0|_3"],
                      initial="|0><0|_3 otimes |0><0|_3"):
    """The dark-state emitter in a lossy cavity truncated at two photons (reference ``:32-53``): Jaynes-Cummings
    coupling on the bright transition, cavity detuned by ``delta_cx``, emitter losses only with ``lindblad``."""
    system_op = ["{}*|2><2|_3 otimes Id_3".format(-delta_xd), " {} * (Id_3 otimes n_3)".format(delta_cx),
                 "{}*(|1><0|_3 otimes b_3 + |0><1|_3 otimes bdagger_3 )".format(hbar * cav_coupl)]
    lindblad_ops = [["|0><1|_3 otimes Id_3", rad_loss]] if lindblad else []
    lindblad_ops.append(["Id_3 otimes b_3", cav_loss])
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=20.48, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="darkmodel_tls_photons", threshold="10", threshold_ratio="0.3", buffer_blocksize="-1",
        dict_zero="16", precision="12", boson_e_max=7, system_op=system_op, pulse_file_x=pulse_file_x,
        pulse_file_y=pulse_file_y, boson_op="|1><1|_3 otimes Id_3 + |2><2|_3 otimes Id_3", initial=initial,
        lindblad_ops=lindblad_ops,
        interaction_ops=[["|2><0|_3 otimes Id_3", "x"], ["|1><2|_3 otimes Id_3", "x"], ["|1><0|_3 otimes Id_3", "y"]],
        output_ops=output_ops, prepare_only=prepare_only)


def _bright_integral(t_end, pulses, n_last, dt, delta_xd, gamma_e, temp_dir, normalize, phonons, pt_file,
                     prepare_only=False):
    t, g, x, d = darkmodel(t_end[0], t_end[1], *pulses, dt=dt, delta_xd=delta_xd,
                           gamma_e=gamma_e, lindblad=True, temp_dir=temp_dir, phonons=phonons, pt_file=pt_file,
                           prepare_only=prepare_only)
    t, x = np.real(t), np.real(x)
    if n_last:
        t, x = t[-n_last:], x[-n_last:]
    val = np.trapezoid(x, t)
    return val / gamma_e if normalize else val


def G1_ee(*pulses, t0=0, dt=0.05, delta_xd=4, gamma_e=1/65, temp_dir=temp_dir, tb=800, normalize=False,
          phonons=False, pt_file=None, prepare_only=False):
    """Bright-exciton occupation integrated over the early bin (reference ``:55-62``)."""
    return _bright_integral((t0, tb), pulses, 0, dt, delta_xd, gamma_e, temp_dir, normalize, phonons, pt_file,
                            prepare_only)


def G1_ll(*pulses, t0=0, dt=0.05, delta_xd=4, gamma_e=1/65, temp_dir=temp_dir, tb=800, normalize=False,
          phonons=False, pt_file=None):
    """... over the late bin: the last ``tb / dt`` rows of a run to ``2 tb`` (reference ``:64-74``)."""
    return _bright_integral((t0, 2 * tb), pulses, int(tb / dt), dt, delta_xd, gamma_e, temp_dir, normalize, phonons,
                            pt_file)


def _el_sweep(pulses, t0, dt, dtau, delta_xd, gamma_e, temp_dir, tb, workers, simple_exp, gaussian_t, phonons,
              pt_file, tend_of, tail):
    """Shared body of :func:`G1_el` / :func:`G1_easy_el` (reference ``:76-129,131-183``): drive files on the ``dtau``
    grid over ``[t0, 2.1 tb)``, one run per ``t1`` with ``|X><G|`` applied from the right -- one GPU batch."""
    from pyaceqd_b200.sweeps import at_time, run_sweep
    if gaussian_t is not None:
        t1 = simple_t_gaussian(t0, gaussian_t, tb, dt, 10 * dt, *pulses)
    else:
        t1 = construct_t(t0, tb, dt, 10 * dt, *pulses, simple_exp=simple_exp)
    grid = np.arange(t0, 2.1 * tb, step=dtau)
    files = []
    for pol in ("x", "y"):
        f = np.zeros_like(grid, dtype=complex)
        for p in pulses:
            f = f + getattr(p, "polar_" + pol) * p.get_total(grid)
        files.append(temp_dir + "G2_pulse_{}.dat".format(pol))
        export_csv(files[-1], grid, f.real, f.imag, precision=8, delimit=' ')
    options = {"dt": dtau, "verbose": False, "delta_xd": delta_xd, "gamma_e": gamma_e, "lindblad": True,
               "pulse_file_x": files[0], "pulse_file_y": files[1], "temp_dir": temp_dir,
               "output_ops": ["|0><0|_3", "|1><1|_3", "|2><2|_3", "|0><1|_3"], "phonons": phonons, "pt_file": pt_file}
    mto = {"operator": "|1><0|_3", "applyFrom": "_right", "applyBefore": "false"}
    jobs = [{"t0": t0, "tend": tend_of(t), "mtos": at_time(mto, t), "tail": tail} for t in t1]
    try:
        res = run_sweep(darkmodel, jobs, *pulses, options=options, workers=workers)
    finally:
        for f in files:
            os.remove(f)
    return t1, res


def G1_el(*pulses, t0=0, dt=0.1, dtau=0.05, delta_xd=4, gamma_e=1/65, temp_dir=temp_dir, tb=800, workers=15,
          normalize=False, simple_exp=False, gaussian_t=None, phonons=False, pt_file=None):
    """``t1, t2, G1[t1, t2]`` (reference ``:76-129``).  As there, EVERY run ends at ``2 tb`` and the rows kept are its
    last ``tb / dtau`` (+ the bright occupation one row earlier), whatever ``t1`` is."""
    n_tau = int(tb / dtau)
    t1, res = _el_sweep(pulses, t0, dt, dtau, delta_xd, gamma_e, temp_dir, tb, workers, simple_exp, gaussian_t,
                        phonons, pt_file, lambda t: 2 * tb, n_tau + 1)
    g1 = np.zeros((len(t1), n_tau + 1), dtype=complex)
    for i, r in enumerate(res):
        g1[i, 0] = r[2][-n_tau - 1]
        g1[i, 1:] = r[4][-n_tau:]
    return t1, np.linspace(0, tb, n_tau + 1), g1


def G1_easy_el(*pulses, t0=0, dt=0.1, dtau=0.05, delta_xd=4, gamma_e=1/65, temp_dir=temp_dir, tb=800, t_offset=0,
               workers=15, normalize=False, simple_exp=False, gaussian_t=None, phonons=False, pt_file=None):
    """``t1, <sigma(t1 + tb + t_offset) sigma^+(t1)>``: only the last value of each run (reference ``:131-183``)."""
    t1, res = _el_sweep(pulses, t0, dt, dtau, delta_xd, gamma_e, temp_dir, tb, workers, simple_exp, gaussian_t,
                        phonons, pt_file, lambda t: t + tb + t_offset, 1)
    return t1, np.array([r[4][-1] for r in res])
