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
