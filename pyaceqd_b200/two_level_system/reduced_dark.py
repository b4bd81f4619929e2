"""Three-level bright/dark exciton model |0>=G, |1>=X, |2>=D (reference
``pyaceqd/two_level_system/reduced_dark.py:13-32``).  The legacy ``G1_el`` sweeps of that file use the
same one-trajectory-per-t idiom as :mod:`pyaceqd_b200.two_time` and are not rebuilt."""
from pyaceqd_b200.general_system.general_system import system_ace_stream
import pyaceqd_b200.constants as constants

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
