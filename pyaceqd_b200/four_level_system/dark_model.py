"""Five-level dark-exciton model ``darkmodel_new`` (reference
``pyaceqd/four_level_system/dark_model.py:34-55``): |0>=G, |1>=X, |2>=Y, |3>=D, |4>=B.  'x'
polarisation drives G-X-B, 'y' drives G-D-B; Y is only reached by decay from B.  The legacy
``G2_*`` workflows of that file are superseded by ``timebin`` and not rebuilt."""
from pyaceqd_b200.general_system.general_system import system_ace_stream
import pyaceqd_b200.constants as constants

hbar = constants.hbar
temp_dir = constants.temp_dir

_POPULATIONS_5 = ["|0><0|_5", "|1><1|_5", "|2><2|_5", "|3><3|_5", "|4><4|_5"]


def darkmodel_new(t_start, t_end, *pulses, dt=0.5, delta_xd=0, delta_b=4, gamma_e=1/100, gamma_b=None,
                  phonons=False, ae=5.0, temperature=4, verbose=False, lindblad=False, temp_dir=temp_dir,
                  pt_file=None, suffix="", multitime_op=None, pulse_file_x=None, pulse_file_y=None,
                  prepare_only=False, threshold=8, output_ops=_POPULATIONS_5, initial="|0><0|_5",
                  use_infinite=True, calc_dynmap=False):
    decay = []
    if lindblad:
        g_b = gamma_e if gamma_b is None else gamma_b
        decay = [["|0><1|_5", gamma_e], ["|0><2|_5", gamma_e], ["|1><4|_5", g_b], ["|2><4|_5", g_b]]
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=20.48, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="darkmodel_new_", threshold=str(int(threshold)), threshold_ratio="0.3",
        buffer_blocksize="-1", dict_zero="16", precision="12", boson_e_max=7,
        system_op=["{}*|4><4|_5".format(-delta_b), "{}*|3><3|_5".format(-delta_xd)],
        pulse_file_x=pulse_file_x, pulse_file_y=pulse_file_y,
        boson_op="1*(|1><1|_5 + |2><2|_5 + |3><3|_5) + 2*|4><4|_5", initial=initial, lindblad_ops=decay,
        interaction_ops=[["|1><0|_5", "x"], ["|4><1|_5", "x"], ["|3><0|_5", "y"], ["|4><3|_5", "y"]],
        output_ops=output_ops, prepare_only=prepare_only, use_infinite=use_infinite, calc_dynmap=calc_dynmap)


def darkmodel(t_start, t_end, *pulses, dt=0.5, delta_xd=0, delta_b=4, gamma_e=1/100, gamma_b=None, phonons=False,
              ae=3.0, temperature=4, verbose=False, lindblad=False, temp_dir=temp_dir, pt_file=None, suffix="",
              multitime_op=None, pulse_file_x=None, pulse_file_y=None, prepare_only=False,
              output_ops=["|0><0|_4", "|1><1|_4", "|2><2|_4", "|3><3|_4"], initial="|0><0|_4"):
    """Four-level dark-exciton model |0>=G, |1>=X, |2>=D, |3>=B (reference ``dark_model.py:13-32``):
    'x' polarisation drives G-D-B, 'y' drives G-X-B; the dark state does not decay."""
    decay = []
    if lindblad:
        decay = [["|0><1|_4", gamma_e], ["|1><3|_4", gamma_e if gamma_b is None else gamma_b]]
    return system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=20.48, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="darkmodel_", threshold="10", threshold_ratio="0.3", buffer_blocksize="-1", dict_zero="16",
        precision="12", boson_e_max=7, system_op=["{}*|3><3|_4".format(-delta_b), "{}*|2><2|_4".format(-delta_xd)],
        pulse_file_x=pulse_file_x, pulse_file_y=pulse_file_y, boson_op="1*(|1><1|_4 + |2><2|_4) + 2*|3><3|_4",
        initial=initial, lindblad_ops=decay,
        interaction_ops=[["|2><0|_4", "x"], ["|3><2|_4", "x"], ["|1><0|_4", "y"], ["|3><1|_4", "y"]],
        output_ops=output_ops, prepare_only=prepare_only)
