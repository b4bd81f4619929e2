"""One-photon time-bin state ``(|e> + |l>)``: populations of the early / late bin and their
coherence from a G1 sweep (reference ``pyaceqd/timebin/onephoton.py:12-106``)."""
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
