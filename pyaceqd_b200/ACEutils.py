"""Stand-in for ACE's pybind module ``ACEutils`` -- exactly the surface the reference touches in its
``calc_dynmap`` / ``get_M_t`` branch (``pyaceqd/general_system/general_system.py:14,313-336``):

    param = Parameters(list_of_param_file_lines)
    initial_state = InitialState(param)  |  InitialState(rho0)
    fprop = FreePropagator(param);  fprop.update(t, dt);  fprop.M
    PT = ProcessTensors(param);  outp = OutputPrinter(param);  tgrid = TimeGrid(param)
    sim = Simulation(param);  sim.run(fprop, PT, initial_state, tgrid, outp)
    DynamicalMap(fprop, PT, sim, tgrid).E

Point ``pyaceqd.constants.pybind_path`` at this package directory and the unmodified reference
imports this file as ``ACEutils``.  Everything runs on the CUDA engine (no CPU fallback):
``Simulation.run`` writes the ``outfile`` the reference then parses (``:334``), ``DynamicalMap.E``
propagates the NL unit vectors in one batch, ``E[i] = E_{t_{i+1}, t_0}`` (``tools.py:470-479``).
"""
from __future__ import annotations

import copy

import numpy as np

from pyaceqd_b200 import ace_cli


def _engine():
    import pyaceqd_b200.engine as eng
    return eng.default_engine()


class Parameters:
    def __init__(self, lines):
        self.lines = [str(l) for l in lines]
        self.parsed = ace_cli.parse_param_lines(self.lines)
        self.problem, self.pt, self.job = ace_cli.setup_from_params(self.parsed)


class InitialState:
    def __init__(self, source):
        if isinstance(source, Parameters):
            self.rho = np.array(source.problem.rho0, dtype=complex)
        else:
            self.rho = np.asarray(source, dtype=complex).reshape(-1)


class FreePropagator:
    """System propagator of one full step: ``update(t, dt)`` evaluates ``M = exp(L(t + dt/2) dt)``."""

    def __init__(self, param: Parameters):
        self._p = param
        self.M = None

    def update(self, t, dt):
        prob, job = self._p.problem, self._p.job
        L = prob.L0.copy()
        for k, pol in enumerate(prob.field_pol):
            tab = job.tables.get(pol)
            if tab is None:
                continue
            x = (t + 0.5 * dt - tab.t0) / tab.dt
            grid = np.arange(len(tab.values))
            f = np.interp(x, grid, tab.values.real) + 1j * np.interp(x, grid, tab.values.imag)
            L = L + f * prob.LA[k] + np.conj(f) * prob.LB[k]
        self.M = _engine().expm(L * dt)[0]


class ProcessTensors:
    def __init__(self, param: Parameters):
        self.pt = param.pt


class OutputPrinter:
    def __init__(self, param: Parameters):
        self.outfile = ace_cli._one(param.parsed, "outfile", "ACE.out")


class TimeGrid:
    def __init__(self, param: Parameters):
        self._job = param.job

    def get_all(self):
        return self._job.times()


class Simulation:
    def __init__(self, param: Parameters):
        self._p = param
        self.initial = None

    def run(self, fprop, PT, initial_state, tgrid, outp):
        p = self._p
        prob = p.problem
        if not np.array_equal(initial_state.rho, prob.rho0):
            prob = copy.copy(prob)
            prob.rho0 = initial_state.rho
        self.initial = initial_state.rho
        out = _engine().run_jobs(prob, PT.pt, [p.job])[0]
        ace_cli.write_outfile(p.parsed, p.job, out)


class DynamicalMap:
    def __init__(self, fprop, PT, sim: Simulation, tgrid):
        p = sim._p
        NL = p.problem.NL
        basis = copy.copy(p.problem)
        basis.out_w = np.eye(NL, dtype=complex)
        jobs = []
        for j in range(NL):
            jb = copy.copy(p.job)
            jb.rho0 = np.zeros(NL, dtype=complex)
            jb.rho0[j] = 1.0
            jobs.append(jb)
        cols = _engine().run_jobs(basis, PT.pt, jobs)
        E = np.empty((p.job.n_steps + 1, NL, NL), dtype=complex)
        for j, c in enumerate(cols):
            E[:, :, j] = c.T
        self.E = [E[i] for i in range(1, len(E))]


def read_outfile(path):
    """Imported (unused) by the reference (``general_system.py:14``)."""
    return np.genfromtxt(path)
