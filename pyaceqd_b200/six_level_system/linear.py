"""Six-level dark/bright exciton system in a magnetic field (reference
``pyaceqd/six_level_system/linear.py:20-72``): |0>=G, |1>=X, |2>=Y, |3>=S(Dx), |4>=F(Dy), |5>=B.
In-plane field ``bx`` mixes bright and dark excitons, ``bz`` mixes within each doublet."""
from pyaceqd_b200.general_system.general_system import system_ace_stream
from pyaceqd_b200.tools import output_ops_dm, compose_dm, read_calibration_file
import pyaceqd_b200.constants as constants

temp_dir = constants.temp_dir
hbar = constants.hbar
d0, d1, d2 = 0.25, 0.12, 0.05   # exchange splittings, meV
mu_b = 5.7882818012e-2          # meV/T
g_ex, g_ez, g_hx, g_hz = -0.65, -0.8, -0.35, -2.2

_POPULATIONS_6 = ["|{0}><{0}|_6".format(i) for i in range(6)]


def energies_linear(d0=0.25, d1=0.12, d2=0.05, delta_B=4, delta_E=0.0):
    """Level energies (meV) relative to ``delta_E`` (reference ``:20-26``)."""
    return (delta_E + (d0 + d1) / 2.0, delta_E + (d0 - d1) / 2.0, delta_E - (d0 - d2) / 2.0,
            delta_E - (d0 + d2) / 2.0, 2. * delta_E - delta_B)


def sixls_linear(t_start, t_end, *pulses, dt=0.5, delta_b=4, gamma_e=1/100, gamma_b=None, gamma_d=0, bx=0, bz=0,
                 phonons=False, ae=3.0, temperature=4, verbose=False, lindblad=False, temp_dir=temp_dir,
                 pt_file=None, suffix="", multitime_op=None, pulse_file_x=None, pulse_file_y=None,
                 prepare_only=False, output_ops=_POPULATIONS_6, initial="|0><0|_6", t_mem=20.48,
                 output_dm=False, dressedstates=False, rf=False, rf_file=None, firstonly=False,
                 calibration_file=None, print_H=False, use_infinite=True, d0=d0, d1=d1, d2=d2):
    if calibration_file is not None:      # measured energies, rates and g factors replace the arguments (:33-34)
        E_X, E_Y, E_S, E_F, E_B, gamma_e, gamma_b, gamma_d, gex, ghx, gez, ghz = read_calibration_file(calibration_file)
    else:
        E_X, E_Y, E_S, E_F, E_B = energies_linear(delta_B=delta_b, d0=d0, d1=d1, d2=d2)
        gex, gez, ghx, ghz = -0.65, -0.8, -0.35, -2.2
    hamiltonian = ["{}*|1><1|_6 + {}*|2><2|_6 + {}*|3><3|_6 + {}*|4><4|_6 + {}*|5><5|_6".format(
        E_X, E_Y, E_S, E_F, E_B)]
    if bx != 0:
        hamiltonian.append("{}*(|1><3|_6 + |3><1|_6 )".format(-0.5 * mu_b * bx * (gex + ghx)))
        hamiltonian.append("{}*(|2><4|_6 + |4><2|_6 )".format(-0.5 * mu_b * bx * (gex - ghx)))
    if bz != 0.0:
        hamiltonian.append("-i*{}*(|2><1|_6 - |1><2|_6 )".format(-0.5 * mu_b * bz * (gez - 3 * ghz)))
        hamiltonian.append("-i*{}*(|4><3|_6 - |3><4|_6 )".format(+0.5 * mu_b * bz * (gez + 3 * ghz)))
    decay = []
    if lindblad:
        g_b = gamma_e if gamma_b is None else gamma_b
        decay = [["|0><1|_6", gamma_e], ["|0><2|_6", gamma_e], ["|1><5|_6", g_b], ["|2><5|_6", g_b],
                 ["|0><3|_6", gamma_d], ["|0><4|_6", gamma_d]]
    if output_dm:
        output_ops = output_ops_dm(dim=6)
    result = system_ace_stream(
        t_start, t_end, *pulses, dt=dt, phonons=phonons, t_mem=t_mem, ae=ae, temperature=temperature,
        verbose=verbose, temp_dir=temp_dir, pt_file=pt_file, suffix=suffix, multitime_op=multitime_op,
        system_prefix="sixls_linear", threshold="10", threshold_ratio="0.3", buffer_blocksize="-1",
        dict_zero="16", precision="12", boson_e_max=7, system_op=hamiltonian, pulse_file_x=pulse_file_x,
        pulse_file_y=pulse_file_y, boson_op="1*(|1><1|_6+|2><2|_6+|3><3|_6+|4><4|_6) + 2*|5><5|_6",
        initial=initial, lindblad_ops=decay, interaction_ops=[["|1><0|_6+|5><1|_6", "x"], ["|2><0|_6+|5><2|_6", "y"]],
        output_ops=output_ops, prepare_only=prepare_only, dressedstates=dressedstates,
        rf_op="|1><1|_6+|2><2|_6+|3><3|_6+|4><4|_6+2*|5><5|_6" if rf else None, rf_file=rf_file,
        firstonly=firstonly, print_H=print_H, use_infinite=use_infinite)
    if output_dm:
        return compose_dm(result, dim=6)
    return result
