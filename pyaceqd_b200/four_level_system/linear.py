"""Biexciton four-level system in the linear-polarisation basis.  Signature and physics of the
reference's ``biexciton`` (``pyaceqd/four_level_system/linear.py:8-39``): |0>=G, |1>=X, |2>=Y,
|3>=B; binding energy ``delta_b``, fine-structure splitting ``delta_xy``; phonon coupling
``|1><1|+|2><2|+2|3><3|`` (coupling classes {0,1,1,2})."""
from pyaceqd_b200.general_system.general_system import system_ace_stream
import pyaceqd_b200.constants as constants

hbar = constants.hbar
temp_dir = constants.temp_dir

_POPULATIONS_4 = ["|0><0|_4", "|1><1|_4", "|2><2|_4", "|3><3|_4"]


def biexciton(t_start, t_end, *pulses, dt=0.5, delta_xy=0, shift_x=True, coupl_xy=0, delta_b=4, gamma_e=1/100,
              gamma_b=None, phonons=False, ae=3.0, temperature=4, verbose=False, lindblad=False,
              temp_dir=temp_dir, pt_file=None, suffix="", multitime_op=None, pulse_file_x=None,
              pulse_file_y=None, prepare_only=False, output_ops=_POPULATIONS_4, initial="|0><0|_4",
              t_mem=20.48, dressedstates=False, rf=False, rf_file=None, firstonly=False, use_infinite=False,
              calc_dynmap=False):
    hamiltonian = ["{}*|3><3|_4".format(-delta_b)]
    if shift_x:   # split X and Y symmetrically around zero
        hamiltonian += ["{}*|1><1|_4".format(-delta_xy / 2), "{}*|2><2|_4".format(delta_xy / 2)]
    else:
        hamiltonian += ["{}*|2><2|_4".format(delta_xy)]
    if coupl_xy != 0:
        hamiltonian += ["{}*|1><2|_4".format(coupl_xy), "{}*|2><1|_4".format(coupl_xy)]
    decay = []
    if lindblad:
        g_b = gamma_e if gamma_b is None else gamma_b
        decay = [["|0><1|_4", gamma_e], ["|0><2|_4", gamma_e], ["|1><3|_4", g_b], ["|2><3|_4", g_b]]
    # the biexciton holds two excitons -> weight 2 in coupling and rotating frame
    two_exciton_weight = "|1><1|_4 + |2><2|_4 + 2*|3><3|_4"
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="b_linear", threshold="10", threshold_ratio="0.3", buffer_blocksize="-1",
        dict_zero="16", precision="12", boson_e_max=7, system_op=hamiltonian, pulse_file_x=pulse_file_x,
        pulse_file_y=pulse_file_y, boson_op="1*(|1><1|_4 + |2><2|_4) + 2*|3><3|_4", initial=initial,
        lindblad_ops=decay, interaction_ops=[["|1><0|_4+|3><1|_4", "x"], ["|2><0|_4+|3><2|_4", "y"]],
        output_ops=output_ops, prepare_only=prepare_only, dressedstates=dressedstates,
        rf_op=two_exciton_weight if rf else None, rf_file=rf_file, firstonly=firstonly,
        use_infinite=use_infinite, calc_dynmap=calc_dynmap)
