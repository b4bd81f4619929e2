"""Two-level quantum dot.  Signature and physics of the reference's ``tls``
(``pyaceqd/two_level_system/tls.py:16-77``): |0> ground, |1> exciton, x-polarised drive on
``|1><0|``, radiative decay ``|0><1|`` (``gamma_e``), optional pure dephasing, phonon coupling
``phonon_factor*|1><1|``; runs on the CUDA engine through ``system_ace_stream``."""
from pyaceqd_b200.general_system.general_system import system_ace_stream
import pyaceqd_b200.constants as constants

hbar = constants.hbar
temp_dir = constants.temp_dir

_TLS_OUTPUTS = ["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2"]


def tls(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, phonons=False, t_mem=6.4, ae=5.0, temperature=4,
        verbose=False, lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None,
        pulse_file=None, pulse_file_x=None, prepare_only=False, output_ops=_TLS_OUTPUTS, phonon_factor=1.0,
        LO_params=None, dressedstates=False, rf=False, rf_file=None, firstonly=False, dephasing=None,
        J_to_file=None, J_file=None, factor_ah=None, use_infinite=True, threshold=8, calc_dynmap=False,
        rho0=None, e_x=0, get_M_t=None, initial="|0><0|_2", **options):
    """Returns ``[t, <out_1>, ...]`` (default outputs: g, x, <|0><1|>, <|1><0|>)."""
    decay = []
    if lindblad:
        decay.append(["|0><1|_2", gamma_e])
    if dephasing is not None:
        decay.append(["|0><0|_2-|1><1|_2", dephasing])
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        pulse_file_x=pulse_file if pulse_file is not None else pulse_file_x, system_prefix="tls",
        threshold=str(int(threshold)), threshold_ratio="0.3", buffer_blocksize="-1", dict_zero="16",
        precision="12", boson_e_max=7,
        system_op=["({}*|1><1|_2)".format(e_x)] if e_x != 0 else None,
        boson_op="{:.3f}*|1><1|_2".format(phonon_factor), initial=initial, lindblad_ops=decay,
        interaction_ops=[["|1><0|_2", "x"]], output_ops=output_ops, prepare_only=prepare_only,
        LO_params=LO_params, dressedstates=dressedstates, rf_op="|1><1|_2" if rf else None, rf_file=rf_file,
        firstonly=firstonly, J_to_file=J_to_file, J_file=J_file, factor_ah=factor_ah,
        use_infinite=use_infinite, calc_dynmap=calc_dynmap, rho0=rho0, get_M_t=get_M_t)


# ------------------------------------------------------------------------------------ composite systems
# The reference's cavity / sensor variants (``tls.py:86-348``) are all "two-level emitter (factor 0)
# (x) bosonic cavities (x) two-level sensors".  One builder writes their operator strings; the variants
# below only differ in the factor list.  The engine handles Liouville dimensions up to 64 (Hilbert
# dimension 8): larger products raise a clear error instead of silently falling back.

def _embed(op, k, dims):
    """``Id (x) ... op at factor k ... (x) Id`` as an ACE operator string."""
    return " otimes ".join(op if i == k else "Id_{}".format(d) for i, d in enumerate(dims))


def _emitter_cavity_sensor(t_start, t_end, pulses, *, prefix, cavities, sensors, sensor_source, dt, gamma_e, lindblad,
                           rf, laser_cav_coupl, output_ops, initial, common):
    """``cavities``: [(levels, detuning, coupling, loss)]; ``sensors``: [(detuning, linewidth)] weakly coupled
    (``epsilon``) to ``sensor_source`` = "emitter" or the index of the cavity they monitor."""
    dims = [2] + [c[0] for c in cavities] + [2] * len(sensors)
    n_hilbert = 1
    for d in dims:
        n_hilbert *= d
    if n_hilbert > 8:
        raise ValueError("{}: Hilbert dimension {} (Liouville {}) exceeds the engine's limit of 8 (64); reduce "
                         "the photon numbers".format(prefix, n_hilbert, n_hilbert ** 2))
    on = lambda op, k: _embed(op, k, dims)
    system_op, decay = [], []
    if lindblad:
        decay.append([on("|0><1|_2", 0), gamma_e])
    drive = [[on("|1><0|_2", 0), "x"]]
    rf_terms = [on("|1><1|_2", 0)]
    for ci, (levels, delta, g, loss) in enumerate(cavities, start=1):
        system_op.append(" {} * ({})".format(delta, on("n_{}".format(levels), ci)))
        jc = lambda a, b: " otimes ".join(a if i == 0 else (b if i == ci else "Id_{}".format(d))
                                          for i, d in enumerate(dims))
        system_op.append(" {} * ({} + {})".format(g, jc("|1><0|_2", "b_{}".format(levels)),
                                                  jc("|0><1|_2", "bdagger_{}".format(levels))))
        decay.append([on("b_{}".format(levels), ci), loss])
        rf_terms.append(on("n_{}".format(levels), ci))
        if laser_cav_coupl is not None and ci == 1:
            drive.append(["{}*({})".format(laser_cav_coupl, on("bdagger_{}".format(levels), ci)), "x"])
    eps = common.pop("epsilon", 0.0001)
    for si, (delta_s, width) in enumerate(sensors):
        k = 1 + len(cavities) + si
        system_op.append("{} * ({})".format(delta_s, on("|1><1|_2", k)))
        if sensor_source == "emitter":
            up, down, src = "|1><0|_2", "|0><1|_2", 0
        else:
            lv = cavities[sensor_source][0]
            up, down, src = "bdagger_{}".format(lv), "b_{}".format(lv), 1 + sensor_source
        pair = lambda a, b: " otimes ".join(a if i == src else (b if i == k else "Id_{}".format(d))
                                            for i, d in enumerate(dims))
        system_op.append("{} * ({} + {})".format(eps, pair(up, "|0><1|_2"), pair(down, "|1><0|_2")))
        decay.append([on("|0><1|_2", k), width])
    if initial is None:
        initial = " otimes ".join("|0><0|_{}".format(d) for d in dims)
    if output_ops is None:
        output_ops = [on("|0><0|_2", 0), on("|1><1|_2", 0)]
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, system_prefix=prefix, threshold="10", threshold_ratio="0.3",
        buffer_blocksize="-1", dict_zero="16", precision="12", boson_e_max=7, system_op=system_op,
        boson_op=on("|1><1|_2", 0), initial=initial, lindblad_ops=decay, interaction_ops=drive, output_ops=output_ops,
        rf_op=" + ".join(rf_terms) if rf else None, **common)


def _common(**kw):
    return {k: v for k, v in kw.items()}


def tls_one_sensor(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, phonons=False, t_mem=10, ae=3.0, delta_s1=0,
                   epsilon=0.0001, linewidth1=0.01, temperature=1, verbose=False, lindblad=False, temp_dir=temp_dir,
                   pt_file=None, suffix="", multitime_op=None, pulse_file=None, prepare_only=False, output_ops=None,
                   initial=None, dressedstates=False, rf=False, rf_file=None, firstonly=False, calc_dynmap=False,
                   use_infinite=False, get_M_t=None):
    """Emitter + one weakly coupled two-level sensor (reference ``tls.py:130-161``)."""
    return _emitter_cavity_sensor(
        t_start, t_end, pulses, prefix="tls_one_sensor", cavities=[], sensors=[(delta_s1, linewidth1)],
        sensor_source="emitter", dt=dt, gamma_e=gamma_e, lindblad=lindblad, rf=rf, laser_cav_coupl=None,
        output_ops=output_ops, initial=initial,
        common=_common(epsilon=epsilon, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature, verbose=verbose,
                       temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
                       pulse_file_x=pulse_file, prepare_only=prepare_only, dressedstates=dressedstates, rf_file=rf_file,
                       firstonly=firstonly, use_infinite=use_infinite, calc_dynmap=calc_dynmap, get_M_t=get_M_t))


def tls_two_sensor(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, phonons=False, t_mem=10, ae=3.0, delta_s1=0,
                   delta_s2=0, epsilon=0.0001, linewidth1=0.01, linewidth2=None, temperature=1, verbose=False,
                   lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None, pulse_file=None,
                   prepare_only=False, output_ops=None, initial=None, dressedstates=False, rf=False, rf_file=None,
                   firstonly=False, calc_dynmap=False, use_infinite=False, get_M_t=None):
    """Emitter + two sensors, e.g. for frequency-filtered correlations (reference ``tls.py:86-128``)."""
    return _emitter_cavity_sensor(
        t_start, t_end, pulses, prefix="tls_two_sensor", cavities=[],
        sensors=[(delta_s1, linewidth1), (delta_s2, linewidth1 if linewidth2 is None else linewidth2)],
        sensor_source="emitter", dt=dt, gamma_e=gamma_e, lindblad=lindblad, rf=rf, laser_cav_coupl=None,
        output_ops=output_ops, initial=initial,
        common=_common(epsilon=epsilon, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature, verbose=verbose,
                       temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
                       pulse_file_x=pulse_file, prepare_only=prepare_only, dressedstates=dressedstates, rf_file=rf_file,
                       firstonly=firstonly, use_infinite=use_infinite, calc_dynmap=calc_dynmap, get_M_t=get_M_t))


def tls_photon(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, cav_coupl1=0.06, cav_loss1=0.12/hbar, delta_cx1=-2,
               phonons=False, t_mem=10, ae=5.0, temperature=4, verbose=False, lindblad=False, temp_dir=temp_dir,
               pt_file=None, suffix="", multitime_op=None, n_phot1=2, laser_cav_coupl=None, pulse_file_x=None,
               prepare_only=False, output_ops=None, dressedstates=False, rf=False, rf_file=None, firstonly=False,
               initial=None, use_infinite=True, calc_dynmap=False, rho0=None, **options):
    """Emitter in one cavity mode with up to ``n_phot1`` photons (Jaynes-Cummings + loss, reference
    ``tls.py:222-257``)."""
    if rf and pulse_file_x is not None and rf_file is None:
        print("Error: pulse file is given, but no file for rotating frame")
        return 0
    return _emitter_cavity_sensor(
        t_start, t_end, pulses, prefix="tls_cavity", cavities=[(n_phot1 + 1, delta_cx1, cav_coupl1, cav_loss1)],
        sensors=[], sensor_source="emitter", dt=dt, gamma_e=gamma_e, lindblad=lindblad, rf=rf,
        laser_cav_coupl=laser_cav_coupl, output_ops=output_ops, initial=initial,
        common=_common(phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature, verbose=verbose, temp_dir=temp_dir,
                       pt_file=pt_file, suffix=suffix, multitime_op=multitime_op, pulse_file_x=pulse_file_x,
                       prepare_only=prepare_only, dressedstates=dressedstates, rf_file=rf_file, firstonly=firstonly,
                       use_infinite=use_infinite, calc_dynmap=calc_dynmap, rho0=rho0))


def tls_photons(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, cav_coupl1=0.06, cav_loss1=0.12/hbar, delta_cx1=-2,
                cav_coupl2=None, cav_loss2=None, delta_cx2=-2, phonons=False, t_mem=10, ae=5.0, temperature=4,
                verbose=False, lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None, n_phot1=2,
                n_phot2=2, laser_cav_coupl=None, pulse_file=None, prepare_only=False, output_ops=None,
                dressedstates=False, rf=False, rf_file=None, firstonly=False, initial=None):
    """Emitter in two cavity modes (reference ``tls.py:163-214``); fits the engine for ``n_phot1 = n_phot2 = 1``."""
    if rf and pulse_file is not None and rf_file is None:
        print("Error: pulse file is given, but no file for rotating frame")
        return 0
    return _emitter_cavity_sensor(
        t_start, t_end, pulses, prefix="tls_cavity",
        cavities=[(n_phot1 + 1, delta_cx1, cav_coupl1, cav_loss1),
                  (n_phot2 + 1, delta_cx2, cav_coupl1 if cav_coupl2 is None else cav_coupl2,
                   cav_loss1 if cav_loss2 is None else cav_loss2)],
        sensors=[], sensor_source="emitter", dt=dt, gamma_e=gamma_e, lindblad=lindblad, rf=rf,
        laser_cav_coupl=laser_cav_coupl, output_ops=output_ops, initial=initial,
        common=_common(phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature, verbose=verbose, temp_dir=temp_dir,
                       pt_file=pt_file, suffix=suffix, multitime_op=multitime_op, pulse_file_x=pulse_file,
                       prepare_only=prepare_only, dressedstates=dressedstates, rf_file=rf_file, firstonly=firstonly))


def tls_photon_sensor(t_start, t_end, *pulses, dt=0.1, gamma_e=1/100, cav_coupl1=0.06, cav_loss1=0.12/hbar, delta_cx1=-2,
                      phonons=False, delta_s1=0, epsilon=0.0001, linewidth1=0.01, t_mem=10, ae=5.0, temperature=4,
                      verbose=False, lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None,
                      n_phot1=2, laser_cav_coupl=None, pulse_file_x=None, prepare_only=False, output_ops=None,
                      dressedstates=False, rf=False, rf_file=None, firstonly=False, initial=None, use_infinite=True,
                      calc_dynmap=False, **options):
    """Emitter + cavity + one sensor on the cavity field (reference ``tls.py:259-300``; ``n_phot1 = 1`` fits)."""
    return _emitter_cavity_sensor(
        t_start, t_end, pulses, prefix="tls_cavity_sensor", cavities=[(n_phot1 + 1, delta_cx1, cav_coupl1, cav_loss1)],
        sensors=[(delta_s1, linewidth1)], sensor_source=0, dt=dt, gamma_e=gamma_e, lindblad=lindblad, rf=rf,
        laser_cav_coupl=laser_cav_coupl, output_ops=output_ops, initial=initial,
        common=_common(epsilon=epsilon, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature, verbose=verbose,
                       temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
                       pulse_file_x=pulse_file_x, prepare_only=prepare_only, dressedstates=dressedstates,
                       rf_file=rf_file, firstonly=firstonly, use_infinite=use_infinite, calc_dynmap=calc_dynmap))


def tls_photon_two_sensor(t_start, t_end, *pulses, n_phot1=2, **kw):
    """Reference ``tls.py:302-348``: emitter (2) x cavity (>= 2) x two sensors (4) has Hilbert dimension >= 16,
    beyond the engine's Liouville-space limit."""
    raise ValueError("tls_photon_two_sensor: Hilbert dimension {} exceeds the engine's limit of 8".format(
        2 * (n_phot1 + 1) * 4))
