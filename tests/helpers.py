"""Shared builders for the parity tests: reference-style problems and jobs."""
import numpy as np

from pyaceqd_b200.jobs import FieldTable, Job
from pyaceqd_b200.problem import build_problem
from pyaceqd_b200.pulses import ChirpedPulse


def tls_problem(lindblad=True, gamma_e=0.01, dephasing=None, e_x=0.0, outputs=None, phonons=True):
    lb = []
    if lindblad:
        lb.append(["|0><1|_2", gamma_e])
    if dephasing:
        lb.append(["|0><0|_2-|1><1|_2", dephasing])
    return build_problem(system_op=["({}*|1><1|_2)".format(e_x)] if e_x else None,
                         boson_op="1.000*|1><1|_2" if phonons else None, initial="|0><0|_2",
                         lindblad_ops=lb, interaction_ops=[["|1><0|_2", "x"]],
                         output_ops=outputs or ["|0><0|_2", "|1><1|_2", "|0><1|_2", "|1><0|_2"])


def biexciton_problem(outputs=None, delta_b=4.0, delta_xy=0.1, lindblad=True, phonons=True):
    lb = []
    if lindblad:
        lb = [["|0><1|_4", 0.01], ["|0><2|_4", 0.01], ["|1><3|_4", 0.012], ["|2><3|_4", 0.012]]
    return build_problem(
        system_op=["{}*|3><3|_4".format(-delta_b), "{}*|1><1|_4".format(-delta_xy / 2), "{}*|2><2|_4".format(delta_xy / 2)],
        boson_op="1*(|1><1|_4 + |2><2|_4) + 2*|3><3|_4" if phonons else None, initial="|0><0|_4", lindblad_ops=lb,
        interaction_ops=[["|1><0|_4+|3><1|_4", "x"], ["|2><0|_4+|3><2|_4", "y"]],
        output_ops=outputs or ["|0><0|_4", "|1><1|_4", "|2><2|_4", "|3><3|_4", "|0><3|_4"])


def sixls_problem(bx=2.0):
    from pyaceqd_b200.six_level_system.linear import energies_linear, mu_b
    E = energies_linear()
    gex, ghx = -0.65, -0.35
    sysop = ["{}*|1><1|_6 + {}*|2><2|_6 + {}*|3><3|_6 + {}*|4><4|_6 + {}*|5><5|_6".format(*E),
             "{}*(|1><3|_6 + |3><1|_6 )".format(-0.5 * mu_b * bx * (gex + ghx)),
             "{}*(|2><4|_6 + |4><2|_6 )".format(-0.5 * mu_b * bx * (gex - ghx))]
    lb = [["|0><1|_6", 0.01], ["|0><2|_6", 0.01], ["|1><5|_6", 0.01], ["|2><5|_6", 0.01]]
    return build_problem(system_op=sysop, boson_op="1*(|1><1|_6+|2><2|_6+|3><3|_6+|4><4|_6) + 2*|5><5|_6",
                         initial="|0><0|_6", lindblad_ops=lb,
                         interaction_ops=[["|1><0|_6+|5><1|_6", "x"], ["|2><0|_6+|5><2|_6", "y"]],
                         output_ops=["|0><0|_6", "|1><1|_6", "|5><5|_6", "|1><0|_6", "|3><3|_6", "|0><5|_6"])


def make_tables(pulses, t_start, t_end, dt, quantise=True):
    t = np.arange(t_start, t_end, dt)
    px = np.zeros_like(t, dtype=complex)
    py = np.zeros_like(t, dtype=complex)
    for p in pulses:
        f = p.get_total(t)
        px = px + p.polar_x * f
        py = py + p.polar_y * f
    if quantise:
        px = np.round(px.real, 8) + 1j * np.round(px.imag, 8)
        py = np.round(py.real, 8) + 1j * np.round(py.imag, 8)
    return {"x": FieldTable(t_start, dt, px), "y": FieldTable(t_start, dt, py)}


def sweep_jobs(n_area, n_det, t_end=8.0, dt=0.1, tau=1.0, t0=4.0):
    """Pulse-area x detuning sweep (scaled)."""
    jobs = []
    for a in np.linspace(0.5, 6.0, n_area):
        for d in np.linspace(-2.0, 2.0, n_det):
            p = ChirpedPulse(tau_0=tau, e_start=d, alpha=0, t0=t0, e0=a)
            jobs.append(Job(0.0, t_end, dt, tables=make_tables([p], 0.0, t_end, dt)))
    return jobs
