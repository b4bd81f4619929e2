"""One-photon time-bin state ``(|e> + |l>)``: populations of the early / late bin and their
coherence from a G1 sweep (reference ``pyaceqd/timebin/onephoton.py:12-106``), and the cavity-filtered
first-order coherences of ``OnePhotonCavity`` (``:108-271``)."""
from __future__ import annotations

import re

import numpy as np

from pyaceqd_b200.sweeps import at_time, run_sweep
from pyaceqd_b200.timebin.timebin import TimeBin
from pyaceqd_b200.tools import construct_t, simple_t_gaussian

options_example = {"verbose": False, "delta_xd": 4, "gamma_e": 1 / 65, "lindblad": True, "temp_dir": "",
                   "phonons": False, "pt_file": None}


class OnePhotonTimebin(TimeBin):
    def __init__(self, system, sigma_x, *pulses, dt=0.02, tb=800, simple_exp=True, gaussian_t=None, verbose=False,
                 workers=15, options={}) -> None:
        super().__init__(system, *pulses, dt=dt, tb=tb, simple_exp=simple_exp, gaussian_t=gaussian_t,
                         verbose=verbose, workers=workers, options=options)
        self.prepare_operators(sigma_x=sigma_x, verbose=verbose)
        if "gamma_e" not in self.options:
            print("gamma_e not supplied in options.")
            exit(1)
        self.gamma_e = self.options["gamma_e"]

    def prepare_operators(self, sigma_x, verbose=False):
        """From the lowering operator ``|l><u|_d`` derive its adjoint and the upper-state projector."""
        m = re.search(r"^\|([0-9]*)><([0-9]*)\|_([1-9]*)", sigma_x)
        lower, upper, dim = m.group(1), m.group(2), m.group(3)
        self.sigma_x = "|{}><{}|_{}".format(lower, upper, dim)
        self.sigma_xdag = "|{}><{}|_{}".format(upper, lower, dim)
        self.x_op = "|{}><{}|_{}".format(upper, upper, dim)
        if verbose:
            print("sigma_x: {}, sigma_xdag: {}, x_op: {}".format(self.sigma_x, self.sigma_xdag, self.x_op))

    def _population_integral(self, tend, last_only, suffix):
        t, x = self.system(0, tend, output_ops=[self.x_op], suffix=suffix, **self.options)
        t, x = np.real(t), np.real(x)
        if last_only:                  # only the late bin counts
            n = int(self.tb / self.dt)
            t, x = t[-n:], x[-n:]
        return np.trapezoid(x, t)

    def rho_ee(self):
        return self._population_integral(self.tb, False, "ee")

    def rho_ll(self):
        return self._population_integral(2 * self.tb, True, "ll")

    def rho_el(self, dt_small=0.1):
        """``G1(t1) = <sigma(t1 + tb) sigma^+(t1)>``: one trajectory per ``t1`` of length ``t1 + tb``,
        only its last output value is used (reference ``:77-106``)."""
        if self.gaussian_t is not None:
            t1 = simple_t_gaussian(0, self.gaussian_t, self.tb, dt_small, 10 * dt_small, *self.pulses)
        else:   # reference quirk kept: first pulse binds to construct_t's positional dt_exp
            t1 = construct_t(0, self.tb, dt_small, 10 * dt_small, *self.pulses, simple_exp=self.simple_exp)
        mto = {"operator": self.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}
        jobs = [{"tend": t + self.tb, "mtos": at_time(mto, t), "output_ops": [self.sigma_x], "tail": 1} for t in t1]
        res = run_sweep(self.system, jobs, options=self.options, workers=self.workers)
        return t1, np.array([r[1][-1] for r in res])

    def calc_densitymatrix(self, first_abs=False, verbose=False):
        """Returns ``rho_ee, rho_ll, |rho_el|, norm`` (reference ``:23-41``); ``first_abs`` integrates
        ``|G1|`` instead of ``G1`` (drops all phase effects)."""
        rho_ee = self.rho_ee() * self.gamma_e
        rho_ll = self.rho_ll() * self.gamma_e
        norm = rho_ee + rho_ll
        t1, g1 = self.rho_el()
        rho_el = (np.trapezoid(np.abs(g1), t1) if first_abs else np.abs(np.trapezoid(g1, t1))) * self.gamma_e
        if verbose:
            print("not normalized:\nEE:{}, LL:{}, EL:{}".format(rho_ee, rho_ll, rho_el))
            print("normalized:\nEE:{}, LL:{}, EL:{}".format(rho_ee / norm, rho_ll / norm, rho_el / norm))
        return rho_ee, rho_ll, rho_el, norm


class OnePhotonCavity(TimeBin):
    """Emitter (3 levels) in a cavity (3 Fock states): ``G1`` of the CAVITY photon, time-integrated over the delay
    (reference ``:108-271``; system: ``two_level_system/reduced_dark.py``).  Every method is one sweep = one GPU
    batch; only the rows a method reads leave the device."""

    def __init__(self, system, *pulses, dt=0.1, tb=20, simple_exp=True, gaussian_t=None, verbose=False, workers=2,
                 t_simul=150, options={}) -> None:
        super().__init__(system, *pulses, dt=dt, tb=tb, simple_exp=simple_exp, gaussian_t=gaussian_t, verbose=verbose,
                         workers=workers, t_simul=t_simul, options=options)
        self.sigma_x = "|0><0|_3 otimes |0><1|_3"
        self.sigma_xdag = "|0><0|_3 otimes |1><0|_3"

    def _axes(self, t0, tend):
        n_t1 = int((tend - t0) / self.dt)
        n_tau = int(self.tb / self.dt)
        return np.linspace(t0, tend, n_t1 + 1), n_tau, np.linspace(-self.tb, self.tb, 2 * n_tau + 1)

    def g1_t1t2(self, t0=30, tend=130, T_sep=0):
        """``int dtau <a^+(t1 - T_sep + tau) a(t1 - T_sep)>`` over ``tau in [-tb, tb]``, the negative delays by
        conjugation (reference ``:115-152``)."""
        t1, n_tau, t2 = self._axes(t0, tend)
        mto = {"operator": self.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}
        outs = ["|0><0|_3 otimes |1><1|_3", self.sigma_x]
        jobs = [{"tend": (t - T_sep) + self.tb, "mtos": at_time(mto, t - T_sep), "output_ops": outs,
                 "tail": n_tau + 1} for t in t1]
        res = run_sweep(self.system, jobs, options=self.options, workers=self.workers)
        g1 = np.zeros(len(t1), dtype=complex)
        for i, r in enumerate(res):
            two_sided = np.zeros(2 * n_tau + 1, dtype=complex)
            two_sided[:n_tau] = np.conjugate(np.flip(r[2][-n_tau:]))
            two_sided[n_tau] = r[1][-(n_tau + 1)]
            two_sided[-n_tau:] = r[2][-n_tau:]
            g1[i] = np.trapezoid(two_sided, t2)
        return t1, g1

    def g1_t1t(self, t0=30, tend=130, T_sep=70):
        """Operator at ``t1 - T_sep``, the last ``2 tb`` of a run to ``t1 + tb`` integrated (reference ``:154-187``)."""
        t1, n_tau, t2 = self._axes(t0, tend)
        mto = {"operator": self.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}
        outs = ["|0><0|_3 otimes |1><1|_3", self.sigma_x]
        n_t2 = 2 * n_tau + 1
        jobs = [{"tend": t + self.tb, "mtos": at_time(mto, t - T_sep), "output_ops": outs, "tail": n_t2} for t in t1]
        res = run_sweep(self.system, jobs, options=self.options, workers=self.workers)
        return t1, np.array([np.trapezoid(r[2][-n_t2:], t2) for r in res])

    def g1_t1(self, t0=30, tend=130, T_sep=70):
        """``G1(t1, t2)`` filled along its anti-diagonals -- a run with the operator at ``t1 + t2 - T_sep`` gives one
        diagonal from its last rows -- and integrated over ``t2`` (reference ``:189-271``).  The reference's three
        executor rounds (rising, full and falling diagonals) are one batch here."""
        t1, n_tau, t2 = self._axes(t0, tend)
        mto = {"operator": self.sigma_x, "applyFrom": "_left", "applyBefore": "false"}
        outs = ["|0><0|_3 otimes |1><1|_3", self.sigma_xdag]
        n_short = min(len(t1), len(t2)) - 1                 # diagonals that do not span the grid
        n_full = len(t1) + len(t2) - 1 - 2 * n_short
        jobs, fill = [], []
        for i in range(n_short):                            # rising: diagonal i has i + 1 elements
            jobs.append({"tend": t1[i], "mtos": at_time(mto, np.round(t1[0] + t2[i] - T_sep, decimals=3)),
                         "output_ops": outs, "tail": i + 1})
            fill.append([(j, i - j, -(i + 1) + j) for j in range(i + 1)])
        for i in range(n_short, n_short + n_full):
            jobs.append({"tend": t1[-1], "mtos": at_time(mto, np.round(t1[i] + t2[0] - T_sep, decimals=3)),
                         "output_ops": outs, "tail": n_short + 1})
            k = i - n_short       # the reference indexes its result list from 0 again and fills rows j + k (:237-241)
            fill.append([(j + k, len(t2) - j - 1, -(n_short + 1) + j) for j in range(n_short + 1)])
        for i in range(n_short):                            # falling
            n_el = n_short - i
            jobs.append({"tend": t1[-1], "mtos": at_time(mto, np.round(t1[i + 1] + t2[-1] - T_sep, decimals=3)),
                         "output_ops": outs, "tail": n_el})
            fill.append([(len(t1) - n_el + j, len(t2) - 1 - j, -n_el + j) for j in range(n_el)])
        res = run_sweep(self.system, jobs, options=self.options, workers=self.workers)
        grid = np.zeros((len(t1), len(t2)), dtype=complex)
        for r, cells in zip(res, fill):
            for (a, b, row) in cells:
                grid[a, b] = r[2][row]
        return t1, np.trapezoid(grid, t2, axis=1)
