"""Single-photon purity and indistinguishability of a pulsed emitter.

``Purity`` and ``Indistinguishability`` keep the constructors, methods and return values of the
reference's ``pyaceqd/two_time/purity.py`` (``Purity`` ``:26-198``, ``Indistinguishability``
``:200-822``): the emitter is driven by a train of identical pulses separated by ``tb``; ``G2(tau)``
and ``|G1(tau)|^2`` integrated over the first bin, evaluated around ``tau = 0`` and around ``tau = tb``,
give purity ``1 - G2[0]/G2[tb]`` and the Hong-Ou-Mandel indistinguishability (``:776-822``).

Two computation routes, as in the reference:

* direct (``dm=False``): one trajectory per ``t`` with operators at ``t`` and ``factor_tau * tb`` of
  further propagation (``:101-140,216-258``) -- here ONE forked GPU batch per sweep;
* time-local maps (``dm=True``, no phonons): the dynamical map of one period is computed once and
  the (t, tau) grid is filled by matrix-vector chains -- the reference's Fortran helper
  (``calc_onetime_parallel_block``, ``:741,770``), here the chain kernel of ``csrc/tlmap.cu``.

With phonons the time-local route (``G1_tl_phonons`` / ``G2_tl_phonons`` ``:513-713``) needs one dynamical-map
run per ``t`` inside the memory time: all of them are ONE GPU batch here, then ``calc_twotime_phonon_block``.
"""
from __future__ import annotations

import numpy as np

from pyaceqd_b200.batch import BatchExecutor, wait
from pyaceqd_b200.pulses import PulseTrain
from pyaceqd_b200.sweeps import at_time, run_sweep
from pyaceqd_b200.timebin.timebin import TimeBin
from pyaceqd_b200.tools import (calc_tl_dynmap_pseudo, construct_t, export_csv, extract_dms, op_to_matrix,
                                simple_t_gaussian)
from pyaceqd_b200.two_time import propagate_tau_module


def _ratio_first_to_second_peak(t, g, tb, dt):
    """``2 * int_0^{tb/2} g`` (the tau = 0 peak is one-sided) and ``int_{tb/2}^{3tb/2} g``."""
    n = int(0.5 * tb / dt)
    return 2 * np.trapezoid(g[:n], t[:n]), np.trapezoid(g[n:3 * n], t[n:3 * n])


class Purity(TimeBin):
    def __init__(self, system, sigma_x, sigma_xdag, *pulses, dt=0.1, tb=800, dt_small=0.1, simple_exp=True,
                 gaussian_t=None, verbose=False, workers=15, t_simul=None, options={}, factor_t=1, factor_tau=2,
                 dt_big=None, add_tend=True) -> None:
        self.factor_t, self.factor_tau = factor_t, factor_tau
        super().__init__(system, PulseTrain(tb, 5, *pulses), dt=dt, tb=tb, simple_exp=simple_exp,
                         gaussian_t=gaussian_t, verbose=verbose, workers=workers, t_simul=t_simul, options=options)
        self.sigma_x, self.sigma_xdag = "(" + sigma_x + ")", "(" + sigma_xdag + ")"
        if "gamma_e" not in options:
            print("gamma_e not included in options, setting to 100")
            self.options["gamma_e"] = 100
        self.gamma_e = self.options["gamma_e"]
        dt_big = 10 * dt_small if dt_big is None else dt_big
        if self.gaussian_t is not None:
            self.t1 = simple_t_gaussian(0, self.gaussian_t, self.tb, dt_small, dt_big, *pulses, decimals=1,
                                        exp_part=self.simple_exp, add_tend=add_tend)
        else:   # reference quirk kept: the first pulse binds to construct_t's positional dt_exp
            self.t1 = construct_t(0, self.tb, dt_small, dt_big, *pulses, simple_exp=self.simple_exp, add_tend=add_tend)
        self.t_axis_complete = np.concatenate([self.t1 + i * self.tb for i in range(factor_t)])
        self.options["pulse_file_x"] = self.pulse_file_x
        self.options["pulse_file_y"] = self.pulse_file_y

    def prepare_pulsefile(self, verbose=False, t_simul=None, plot=False):
        """The pulse train on the simulation grid up to ``(factor_t + factor_tau + 1) tb`` (``:69-91``)."""
        t_end = (self.factor_t + self.factor_tau + 1) * self.tb if t_simul is None else t_simul
        grid = np.linspace(0, t_end, int(t_end / self.dt) + 1)
        fx, fy = self.pulses[0].get_total_xy(grid)
        self.pulse_file_x = self._write("twotime_pulse_x_{}.dat", grid, fx, verbose)
        self.pulse_file_y = self._write("twotime_pulse_y_{}.dat", grid, fy, verbose)

    def calc_timedynamics(self, output_ops=None, t_end=None):
        opts = dict(self.options)
        if output_ops is not None:
            opts["output_ops"] = output_ops
        if t_end is None:
            t_end = (self.factor_t + self.factor_tau + 1) * self.tb
        return self.system(0, t_end, *self.pulses, **opts)

    # ------------------------------------------------------------------ direct route
    def _n_tau(self):
        return self.factor_tau * int(self.tb / self.dt)

    def _grid(self, mtos, output_ops):
        """``[len(t_axis_complete), n_tau + 1]`` complex: tau = 0 from the second output at the operator
        time, tau > 0 from the first output over the last ``n_tau`` rows."""
        n_tau = self._n_tau()
        jobs = [{"tend": t + self.factor_tau * self.tb, "mtos": [at_time(m, t) for m in mtos],
                 "output_ops": output_ops, "tail": n_tau + 1} for t in self.t_axis_complete]
        res = run_sweep(self.system, jobs, options=self.options, workers=self.workers)
        grid = np.empty((len(jobs), n_tau + 1), dtype=complex)
        for i, r in enumerate(res):
            grid[i, 0] = r[2][-(n_tau + 1)]
            grid[i, 1:] = r[1][-n_tau:]
        return np.linspace(0, self.factor_tau * self.tb, n_tau + 1), grid

    def G2_modified(self, out_op1, return_whole=False, tqdm_options={}):
        """``<sigma^+(t) B(t + tau) sigma(t)>`` for a chosen ``B`` (reference ``:142-189``)."""
        mtos = [{"operator": self.sigma_x, "applyFrom": "_left", "applyBefore": "false"},
                {"operator": self.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}]
        t2, grid = self._grid(mtos, [out_op1, self.sigma_xdag + "*" + out_op1 + "*" + self.sigma_x])
        grid = np.abs(grid)
        if return_whole:
            return self.t1, t2, grid
        return t2, np.trapezoid(grid, self.t_axis_complete, axis=0)

    def G2(self, return_whole=False, tqdm_options={}):
        """``G2(tau) = int dt <sigma^+(t) sigma^+ sigma(t + tau) sigma(t)>`` (reference ``:101-140``)."""
        return self.G2_modified(self.sigma_xdag + "*" + self.sigma_x, return_whole=return_whole)

    def calc_purity(self):
        t, g2 = self.G2()
        a, b = _ratio_first_to_second_peak(t, g2, self.tb, self.dt)
        return 1 - a / b


class Indistinguishability(Purity):
    def __init__(self, system, sigma_x, sigma_xdag, *pulses, dt=0.1, tb=800, dt_small=0.1, simple_exp=True,
                 gaussian_t=None, verbose=False, workers=15, t_simul=None, options={}, dm=False, sigma_x_mat=None,
                 sigma_xdag_mat=None, t_mem=10, dt_big=None, add_tend=True) -> None:
        self.dm = dm
        self.tl_map = self.tl_dms = None
        self.t_mem = t_mem
        if sigma_x_mat is None or sigma_xdag_mat is None:
            print("WARNING: sigma_x_mat or sigma_xdag_mat not provided, trying to convert sigma_x and sigma_xdag to matrices")
            sigma_x_mat, sigma_xdag_mat = op_to_matrix(sigma_x), op_to_matrix(sigma_xdag)
        self.sigma_x_mat, self.sigma_xdag_mat = np.asarray(sigma_x_mat), np.asarray(sigma_xdag_mat)
        self.dim = self.sigma_x_mat.shape[0]
        super().__init__(system, sigma_x, sigma_xdag, *pulses, dt=dt, tb=tb, dt_small=dt_small, simple_exp=simple_exp,
                         gaussian_t=gaussian_t, verbose=verbose, workers=workers, t_simul=t_simul, options=options,
                         dt_big=dt_big, add_tend=add_tend)

    # ------------------------------------------------------------------ direct route
    def G1(self):
        """``int dt |<sigma^+(t + tau) sigma(t)>|^2`` (reference ``:216-258``)."""
        t2, grid = self._grid([{"operator": self.sigma_x, "applyFrom": "_left", "applyBefore": "false"}],
                              [self.sigma_xdag, self.sigma_xdag + "*" + self.sigma_x])
        return t2, np.trapezoid(np.abs(grid) ** 2, self.t_axis_complete, axis=0)

    @staticmethod
    def _uncorrelated(val, n_t1, dt_axis_t1, n_t2):
        """``G0(tau) = int dt n(t) n(t + tau)`` from the occupation on the simulation grid (``:281-292``)."""
        out = np.zeros(n_t2)
        for j in range(n_t2):
            shifted = val[j:j + n_t1]
            out[j] = np.trapezoid(val[:len(shifted)] * shifted, dt_axis_t1[:len(shifted)])
        return out

    def simple_propagation(self, return_whole=False):
        n_tau = self._n_tau()
        t2 = np.linspace(0, self.factor_tau * self.tb, n_tau + 1)
        t, val = self.system(0, (self.factor_t + self.factor_tau) * self.tb, suffix=-1,
                             output_ops=[self.sigma_xdag + "*" + self.sigma_x], **self.options)
        t1 = np.linspace(0, self.factor_t * self.tb, int(self.factor_t * self.tb / self.dt) + 1)
        return t2, self._uncorrelated(np.abs(val), len(t1), t1, len(t2))

    # ------------------------------------------------------------------ time-local route (no phonons)
    def get_tl(self, t_mem=None):
        """Dynamical map of one excitation: explicit maps for the first ``gaussian_t`` (or ``tb``), then the
        stationary map (reference ``:395-413``)."""
        if t_mem is None:
            t_mem = self.gaussian_t if self.gaussian_t is not None else self.tb / 2
        result, dm = self.system(0, 2 * t_mem, multitime_op=[], calc_dynmap=True, **self.options)
        t = np.round(result[0].real, 6)
        memory = self.gaussian_t if self.gaussian_t is not None else self.tb
        self.tl_map, pieces = extract_dms(calc_tl_dynmap_pseudo(dm, t), t, memory, t_MTOs=[])
        self.tl_dms = pieces[0]

    def calc_timedynamics_tl(self):
        """Density matrix over ``factor_t + factor_tau`` periods from the per-period maps (``:449-473``)."""
        if self.tl_map is None:
            self.get_tl()
        periods = self.factor_t + self.factor_tau
        n_tb = int(self.tb / self.dt)
        t_total = np.linspace(0, periods * self.tb, periods * n_tb + 1)
        NL = self.dim ** 2
        rho = np.zeros((len(t_total), NL), dtype=complex)
        rho[0, 0] = 1.0
        self.tl_complete = np.zeros((len(t_total) - 1, NL, NL), dtype=complex)
        n_explicit = len(self.tl_dms)
        for k in range(len(t_total) - 1):
            i = k % n_tb                       # the reference uses explicit map i for i < len(tl_dms) - 1
            m = self.tl_dms[i] if i < n_explicit - 1 else self.tl_map
            self.tl_complete[k] = m
            rho[k + 1] = m @ rho[k]
        return t_total, rho.reshape(len(t_total), self.dim, self.dim)

    def simple_propagation_tl(self, return_whole=False):
        t_total, rho = self.calc_timedynamics_tl()
        n_tau = self._n_tau()
        t2 = np.linspace(0, self.factor_tau * self.tb, n_tau + 1)
        t1 = np.linspace(0, self.factor_t * self.tb, int(self.factor_t * self.tb / self.dt) + 1)
        op = self.sigma_xdag_mat @ self.sigma_x_mat
        val = np.real(np.einsum("ab,tba->t", op, rho))
        return t2, self._uncorrelated(val, len(t1), t1, len(t2))

    def _tl_grid(self, opa, opb, opc):
        if self.tl_map is None:
            self.get_tl()
        tau_max = self.tb * self.factor_tau
        n_tau = int(tau_max / self.dt)
        t_end = self.t_axis_complete[-1] + tau_max
        t_axis = np.linspace(0, t_end, int(t_end / self.dt) + 1)
        rho0 = np.zeros(self.dim ** 2, dtype=complex)
        rho0[0] = 1.0
        grid = propagate_tau_module.calc_onetime_parallel_block(
            dm_block=np.asfortranarray(np.asarray(self.tl_dms).transpose(1, 2, 0)), dm_s=self.tl_map, rho_init=rho0,
            n_tb=int(self.tb / self.dt), nx_tau=self.factor_tau, dim=self.dim, opa=opa, opb=opb, opc=opc,
            time=t_axis, time_sparse=self.t_axis_complete)
        return np.linspace(0, tau_max, n_tau + 1), grid

    def G2_tl(self):
        """Reference ``:715-745``."""
        a, c = self.sigma_xdag_mat, self.sigma_x_mat
        tau, grid = self._tl_grid(a, a @ c, c)
        return tau, np.trapezoid(np.abs(grid), self.t_axis_complete, axis=0)

    def G1_tl(self):
        """Reference ``:747-774``."""
        tau, grid = self._tl_grid(np.identity(self.dim), self.sigma_xdag_mat, self.sigma_x_mat)
        return tau, np.trapezoid(np.abs(grid) ** 2, self.t_axis_complete, axis=0)

    # ------------------------------------------------------------------ time-local route with phonons
    def get_tl_phonons(self, mtos=[], t_mtos=[]):
        """Stationary map and explicit blocks of ``gaussian_t + t_mem`` length at the start and after every
        operator time of one run over ``2.1 (gaussian_t + t_mem)`` (reference ``:415-424``)."""
        tmem = self.gaussian_t + self.t_mem
        result, dm = self.system(0, 2.1 * tmem, multitime_op=mtos, calc_dynmap=True, **self.options)
        t = np.round(result[0].real, 6)
        tl_map, pieces = extract_dms(calc_tl_dynmap_pseudo(dm, t), t, tmem, t_MTOs=t_mtos)
        if any(len(b) != len(pieces[0]) for b in pieces):
            raise ValueError("the run over 2.1 (gaussian_t + t_mem) ends before the memory after the operators does: "
                             "0.1 (gaussian_t + t_mem) must cover 6 time steps")
        return tl_map, np.array(pieces, dtype=complex)

    def _periodic_states(self, block, tl_map):
        periods = self.factor_t + self.factor_tau
        n_tb = int(self.tb / self.dt)
        t_total = np.linspace(0, periods * self.tb, periods * n_tb + 1)
        rho = np.zeros((len(t_total), self.dim ** 2), dtype=complex)
        rho[0, 0] = 1.0
        n_explicit = len(block)
        for k in range(len(t_total) - 1):
            i = k % n_tb                       # explicit map i for i < len(block) - 1, as in the reference loops
            rho[k + 1] = (block[i] if i < n_explicit - 1 else tl_map) @ rho[k]
        return t_total, rho

    def calc_timedynamics_tl_phonons(self):
        """Reference ``:426-447``."""
        tl_map, pieces = self.get_tl_phonons(mtos=[], t_mtos=[])
        t_total, rho = self._periodic_states(pieces[0], tl_map)
        return t_total, rho.reshape(len(t_total), self.dim, self.dim)

    def simple_propagation_tl_phonons(self, return_whole=False):
        """``G0(tau)`` from the per-period maps of the phonon run (reference ``:350-393``)."""
        tl_map, pieces = self.get_tl_phonons(mtos=[], t_mtos=[])
        _, rho = self._periodic_states(pieces[0], tl_map)
        rho = rho.reshape(-1, self.dim, self.dim)
        n_tau = self._n_tau()
        t2 = np.linspace(0, self.factor_tau * self.tb, n_tau + 1)
        t1 = np.linspace(0, self.factor_t * self.tb, int(self.factor_t * self.tb / self.dt) + 1)
        val = np.real(np.einsum("ab,tba->t", self.sigma_xdag_mat @ self.sigma_x_mat, rho))
        return t2, self._uncorrelated(val, len(t1), t1, len(t2))

    def _moved(self, mtos, t_mto):
        out = []
        for m in mtos:
            m = dict(m)
            m["time"] = t_mto
            out.append(m)
        return out

    def get_dm2_phonons(self, mtos, t_mto, suffix=1):
        """Explicit maps of the ``gaussian_t + t_mem`` after operators applied at ``t_mto`` (reference ``:475-486``)."""
        tmem = self.gaussian_t + self.t_mem
        result, dm = self.system(0, t_mto + tmem + 2 * self.dt, multitime_op=self._moved(mtos, t_mto),
                                 calc_dynmap=True, suffix=suffix, **self.options)
        t = np.round(result[0].real, 6)
        return extract_dms(calc_tl_dynmap_pseudo(dm, t), t, tmem, t_MTOs=[t_mto])[1][1]

    def _dm2_advanced_run(self, submit, mtos, t_mto, suffix):
        return submit(self.system, 0, self.gaussian_t + 2 * self.t_mem + 2 * self.dt,
                      multitime_op=self._moved(mtos, t_mto), calc_dynmap=True, suffix=suffix, **self.options)

    def _dm2_advanced_maps(self, run, t_mto):
        result, dm = run
        t = np.round(result[0].real, 6)
        memory = np.max([self.gaussian_t + self.t_mem - t_mto, self.t_mem])
        return extract_dms(calc_tl_dynmap_pseudo(dm, t), t, memory, t_MTOs=[t_mto])[1][1]

    def get_dm2_phonons_advanced(self, mtos, t_mto, suffix=1):
        """As :meth:`get_dm2_phonons` with the shortest run that still covers the memory: it ends at ``gaussian_t +
        2 t_mem`` and keeps ``max(gaussian_t + t_mem - t_mto, t_mem)`` of maps (reference ``:488-511``)."""
        return self._dm2_advanced_maps(self._dm2_advanced_run(lambda f, *a, **k: f(*a, **k), mtos, t_mto, suffix), t_mto)

    def _tl_phonon_grid(self, mtos, opa, opb, opc):
        """Shared body of :meth:`G1_tl_phonons` / :meth:`G2_tl_phonons` (reference ``:513-644,646-712``): the
        stationary blocks from a run with the operators beyond all memory, one dynamical-map run per ``t1`` inside
        the memory -- ALL of them in one GPU batch -- and the (t, tau) grid as chains on the chain kernel."""
        t_apply = self.gaussian_t + self.t_mem + 5 * self.dt
        tl_map, pieces = self.get_tl_phonons(mtos=self._moved(mtos, t_apply), t_mtos=[np.round(t_apply, 6)])
        tau_max = self.tb * self.factor_tau
        n_tau = int(tau_max / self.dt)
        inside = np.where(self.t1 <= (self.gaussian_t + self.t_mem))[0]
        own = np.zeros((len(inside),) + pieces[0].shape, dtype=complex)
        own[:, :] = tl_map                       # shorter blocks are padded with the stationary map (:527-529)
        with BatchExecutor(max_workers=self.workers) as ex:
            runs = [self._dm2_advanced_run(ex.submit, mtos, np.round(self.t1[i], 6), int(i)) for i in range(len(inside))]
            wait(runs)
        for i, f in enumerate(runs):
            part = self._dm2_advanced_maps(f.result(), np.round(self.t1[i], 6))
            own[i, :len(part)] = part
        t_end = self.t_axis_complete[-1] + tau_max
        t_axis = np.linspace(0, t_end, int(t_end / self.dt) + 1)
        rho0 = np.zeros(self.dim ** 2, dtype=complex)
        rho0[0] = 1.0
        grid = propagate_tau_module.calc_twotime_phonon_block(
            dm_taucs2=np.asfortranarray(own.transpose(2, 3, 0, 1)),
            dm_sep1=np.asfortranarray(pieces[0].transpose(1, 2, 0)),
            dm_sep2=np.asfortranarray(pieces[1].transpose(1, 2, 0)), dm_s=tl_map, rho_init=rho0,
            n_tb=int(self.tb / self.dt), nx_tau=self.factor_tau, dim=self.dim, opa=opa, opb=opb, opc=opc,
            time=t_axis, time_sparse=self.t_axis_complete)
        return np.linspace(0, tau_max, n_tau + 1), grid

    def G1_tl_phonons(self):
        """Reference ``:513-644``."""
        mtos = [{"operator": self.sigma_x, "applyFrom": "_left", "applyBefore": "false"}]
        tau, grid = self._tl_phonon_grid(mtos, np.identity(self.dim), self.sigma_xdag_mat, self.sigma_x_mat)
        return tau, np.trapezoid(np.abs(grid) ** 2, self.t_axis_complete, axis=0)

    def G2_tl_phonons(self):
        """Reference ``:646-712``."""
        mtos = [{"operator": self.sigma_x, "applyFrom": "_left", "applyBefore": "false"},
                {"operator": self.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}]
        a, c = self.sigma_xdag_mat, self.sigma_x_mat
        tau, grid = self._tl_phonon_grid(mtos, a, a @ c, c)
        return tau, np.trapezoid(np.abs(grid), self.t_axis_complete, axis=0)

    def calc_indistinguishability(self):
        """Returns ``(indistinguishability, single-photon purity)`` (reference ``:776-822``)."""
        phonons = bool(self.dm and self.options.get("phonons"))
        if phonons:
            t1, g1 = self.G1_tl_phonons()
            t2, g2 = self.G2_tl_phonons()
            t0, g0 = self.simple_propagation_tl_phonons()
        else:
            t1, g1 = self.G1_tl() if self.dm else self.G1()
            t2, g2 = self.G2_tl() if self.dm else self.G2()
            t0, g0 = self.simple_propagation_tl() if self.dm else self.simple_propagation()
        g11, g12 = _ratio_first_to_second_peak(t1, g1, self.tb, self.dt)
        g21, g22 = _ratio_first_to_second_peak(t2, g2, self.tb, self.dt)
        g01, g02 = _ratio_first_to_second_peak(t0, g0, self.tb, self.dt)
        return 1 - (g01 - g11 + g21) / (g02 - g12 + g22), 1 - g21 / g22
