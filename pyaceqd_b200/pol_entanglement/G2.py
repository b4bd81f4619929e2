"""Polarisation entanglement of the biexciton-exciton cascade from G2 correlation functions.

Class name (including its spelling), constructor arguments, methods and return layout follow the
reference's ``pyaceqd/pol_entanglement/G2.py`` (``PolarizatzionEntanglement`` ``:11-606``).  The
two-photon density matrix in the basis ``|xx>, |xy>, |yx>, |yy>`` is assembled from ten
time-integrated ``G2`` functions (``:124-159``) or from three sweeps whose trajectories are reused
for several output operators (``:301-356,439-533``).  Each sweep over ``t1`` is ONE GPU batch here
(trunk/branch forking at the operator time) instead of ``len(t1)`` ACE subprocesses.
"""
from __future__ import annotations

import os

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.sweeps import at_time, run_sweep, symmetrised_spectrum, tail_series
from pyaceqd_b200.tools import concurrence, construct_t, export_csv, simple_t_gaussian

hbar = constants.hbar
temp_dir = constants.temp_dir

# (row, col) of the two-photon density matrix <- (sweep, component) of calc_densitymatrix_reuse
_DM_BASIS = ("xx", "xy", "yx", "yy")


def _paren(op: str) -> str:
    return "(" + op + ")"


class PolarizatzionEntanglement():
    def __init__(self, system, sigma_x, sigma_y, sigma_xdag, sigma_ydag, *pulses, dt=0.1, tend=400,
                 time_intervals=None, simple_exp=True, dt_small=0.1, gaussian_t=None, regular_grid=False,
                 verbose=False, workers=2, remove_files=True, factor_tau=4, options={}) -> None:
        """``system``: adapter (e.g. ``sixls_linear``); ``sigma_*``: polarisation lowering / raising
        operator strings; ``tend``: length of the detection window; the ``t1`` grid is regular
        (``regular_grid``), piecewise (``time_intervals``), adaptive (``gaussian_t``) or pulse-refined
        (default), exactly as in the reference (``:84-103``)."""
        self.system = system
        self.dt = dt
        self.options = dict(options)
        self.options["dt"] = dt
        self.tend = tend
        self.remove_files = remove_files
        self.simple_exp = simple_exp
        self.gaussian_t = gaussian_t
        self.pulses = pulses
        self.workers = workers
        self.ax, self.ay = _paren(sigma_x), _paren(sigma_y)
        self.axdag, self.aydag = _paren(sigma_xdag), _paren(sigma_ydag)
        if "temp_dir" not in options:
            print("temp_dir not included in options, setting to temp_dir specified in constants")
            self.options["temp_dir"] = temp_dir
        self.temp_dir = self.options["temp_dir"]
        self.pulse_file_x = self.pulse_file_y = None
        if self.options.get("pulse_file_x") is not None and self.options.get("pulse_file_y") is not None:
            self.remove_files = False
        else:
            self.prepare_pulsefile(verbose=verbose)
            self.options["pulse_file_x"] = self.pulse_file_x
            self.options["pulse_file_y"] = self.pulse_file_y
        self.gamma_e = options["gamma_e"]

        if regular_grid:
            self.t1 = np.arange(0, self.tend + dt_small, dt_small)
        elif time_intervals is not None:
            if len(time_intervals) != 2:
                raise ValueError("time_intervals must be a list of length 2")
            a, b = time_intervals
            self.t1 = np.concatenate([np.arange(0, a, dt_small), np.arange(a, b, 10 * dt_small),
                                      np.round(np.exp(np.arange(np.log(b), np.log(tend), dt_small))),
                                      np.array([tend])])
        elif self.gaussian_t is not None:
            self.t1 = simple_t_gaussian(0, self.gaussian_t, self.tend, dt_small, 10 * dt_small, *self.pulses,
                                        decimals=1, exp_part=self.simple_exp)
        else:
            self.t1 = construct_t(0, self.tend, dt_small, 1 * dt_small, dt_small, *self.pulses,
                                  simple_exp=self.simple_exp, factor_tau=factor_tau)

    # ------------------------------------------------------------------ drive
    def prepare_pulsefile(self, verbose=False):
        """x / y drive sampled on ``dt/5`` over the detection window (reference ``:105-117``)."""
        grid = np.arange(0, self.tend, step=self.dt / 5)
        self.pulse_file_x = self.temp_dir + "polar_ent_pulse_x_{}.dat".format(id(self))
        self.pulse_file_y = self.temp_dir + "polar_ent_pulse_y_{}.dat".format(id(self))
        fx = np.zeros_like(grid, dtype=complex)
        fy = np.zeros_like(grid, dtype=complex)
        for p in self.pulses:
            f = p.get_total(grid)
            fx, fy = fx + p.polar_x * f, fy + p.polar_y * f
        export_csv(self.pulse_file_x, grid, fx.real, fx.imag, precision=8, delimit=' ', verbose=verbose)
        export_csv(self.pulse_file_y, grid, fy.real, fy.imag, precision=8, delimit=' ', verbose=verbose)

    def __del__(self):
        if getattr(self, "remove_files", False):
            for f in (self.pulse_file_x, self.pulse_file_y):
                if f is not None and os.path.exists(f):
                    os.remove(f)

    # ------------------------------------------------------------------ sweeps
    def _n_tau(self):
        return int(self.tend / self.dt)

    def _sweep(self, mtos, output_ops, tend_of, tail_of, tail_reduce=None):
        jobs = [{"tend": tend_of(t), "mtos": [at_time(m, t) for m in mtos], "output_ops": output_ops,
                 "tail": tail_of(t)} for t in self.t1]
        return run_sweep(self.system, jobs, options=self.options, workers=self.workers, tail_reduce=tail_reduce)

    def calc_timedynamics(self, output_ops=None):
        opts = dict(self.options)
        if output_ops is not None:
            opts["output_ops"] = output_ops
        return self.system(0, self.tend, **opts)

    def G1(self, op1_t, op2_ttau):
        """``<op2(t1 + tau) op1(t1)>`` on ``t1 x [0, tend]`` (reference ``:161-205``)."""
        if op1_t[0] != "(":
            op1_t = _paren(op1_t)
            print("WARNING: added brackets to op1_t")
        if op2_ttau[0] != "(":
            op2_ttau = _paren(op2_ttau)
            print("WARNING: added brackets to op2_ttau")
        n_tau = self._n_tau()
        tau = np.linspace(0, self.tend, n_tau + 1)
        res = self._sweep([{"operator": op1_t, "applyFrom": "_left", "applyBefore": "false"}],
                          [op2_ttau, op2_ttau + " * " + op1_t], lambda t: t + self.tend, lambda t: n_tau + 1)
        g1 = np.array([tail_series(r, n_tau) for r in res])
        return self.t1, tau, g1

    def get_spectrum(self, op1_t, op2_ttau, save_g1_dir=None, load=None):
        """Spectrum of ``G1`` (reference ``:207-241``); optional caching of G1 as ``.npy``."""
        if load is not None and os.path.exists(load + "g1.npy"):
            t_axis, tau_axis, g1 = (np.load(load + n) for n in ("t_axis.npy", "tau_axis.npy", "g1.npy"))
        else:
            t_axis, tau_axis, g1 = self.G1(op1_t, op2_ttau)
        if save_g1_dir is not None and load is None:
            np.save(save_g1_dir + "g1.npy", g1)
            np.save(save_g1_dir + "t_axis.npy", t_axis)
            np.save(save_g1_dir + "tau_axis.npy", tau_axis)
        return symmetrised_spectrum(t_axis, tau_axis, g1, hbar)

    def _window_series(self, res, n_ops):
        """Per ``t1``: ``G2(t1, tau)`` for ``tau`` in ``[0, tend - t1]`` of every operator pair
        (``n_t2 = n_tau - int(t1/dt)``: the float truncation is part of the contract, ``:283``)."""
        n_tau = self._n_tau()
        out = []
        for t, r in zip(self.t1, res):
            n_t2 = n_tau - int(t / self.dt)
            out.append(np.array([tail_series(r, n_t2, i_tau=1 + j, i_zero=1 + n_ops + j) for j in range(n_ops)]))
        return out

    def G2(self, op1_t, op2_ttau, op3_ttau, op4_t):
        """``<op1(t1) op2(t1+tau) op3(t1+tau) op4(t1)>`` integrated over ``tau`` (per ``t1``) and over
        both (reference ``:243-299``).  Returns ``t1, G2(t1), integral``."""
        t1, g2, total = self.G2_reuse(op1_t, [op2_ttau + " * " + op3_ttau], op4_t)
        return t1, g2[0], total[0]

    def G2_reuse(self, op1_t, op23s_ttau, op4_t, return_full_G2=False):
        """One sweep, several ``(op2 op3)`` output pairs (reference ``:439-533``)."""
        n_ops = len(op23s_ttau)
        n_tau = self._n_tau()
        tau = np.linspace(0, self.tend, n_tau + 1)
        outputs = list(op23s_ttau) + [op1_t + " * " + o + " * " + op4_t for o in op23s_ttau]
        mtos = [{"operator": op1_t, "applyFrom": "_right", "applyBefore": "false"},
                {"operator": op4_t, "applyFrom": "_left", "applyBefore": "false"}]
        if not return_full_G2:
            # the tau integral of every t1 is taken where the trajectories end, on the device (reference :525-527 does
            # np.trapz over the last n_t2 + 1 rows of every run on the host): only n_t1 x n_ops numbers come back
            red = self._sweep(mtos, outputs, lambda t: self.tend, lambda t: n_tau - int(t / self.dt) + 1,
                              tail_reduce=([(j, n_ops + j) for j in range(n_ops)], tau[1] - tau[0] if n_tau else 0.0))
            g2 = np.array(red, dtype=complex).T.reshape(n_ops, len(self.t1))
            return self.t1, g2, np.trapezoid(g2, self.t1, axis=1)
        res = self._sweep(mtos, outputs, lambda t: self.tend, lambda t: n_tau - int(t / self.dt) + 1)
        series = self._window_series(res, n_ops)
        g2 = np.zeros((n_ops, len(self.t1)), dtype=complex)
        full = np.zeros((n_ops, len(self.t1), n_tau + 1), dtype=complex) if return_full_G2 else None
        for i, s in enumerate(series):
            g2[:, i] = np.trapezoid(s, tau[:s.shape[1]], axis=1)
            if full is not None:
                full[:, i, :s.shape[1]] = s
        total = np.trapezoid(g2, self.t1, axis=1)
        if return_full_G2:
            return self.t1, tau, g2, total, full
        return self.t1, g2, total

    # ------------------------------------------------------------------ two-photon density matrix
    def calc_densitymatrix(self):
        """Ten separate G2 sweeps (reference ``:124-159``); returns the concurrence."""
        a = {"x": (self.ax, self.axdag), "y": (self.ay, self.aydag)}
        rho = np.zeros((4, 4), dtype=complex)
        for r, bra in enumerate(_DM_BASIS):
            for c, ket in enumerate(_DM_BASIS):
                if c < r:
                    continue
                # <bra| .. |ket>:  first photon bra[0]/ket[0] at t1, second photon bra[1]/ket[1] at t1+tau
                _, _, rho[r, c] = self.G2(a[bra[0]][1], a[bra[1]][1], a[ket[1]][0], a[ket[0]][0])
                rho[c, r] = np.conj(rho[r, c])
        return concurrence(rho / np.trace(rho))

    def _reuse_sweeps(self, full=False):
        xx, xy = self.axdag + " * " + self.ax, self.axdag + " * " + self.ay
        yx, yy = self.aydag + " * " + self.ax, self.aydag + " * " + self.ay
        plan = [(self.axdag, [xx, xy, yy], self.ax),          # xx,xx  xx,xy  xy,xy
                (self.axdag, [xx, xy, yx, yy], self.ay),      # xx,yx  xx,yy  xy,yx  xy,yy
                (self.aydag, [xx, xy, yy], self.ay)]          # yx,yx  yx,yy  yy,yy
        return [self.G2_reuse(o1, o23, o4, return_full_G2=full) for (o1, o23, o4) in plan]

    @staticmethod
    def _assemble(c1, c2, c3):
        """4x4 matrix (or a stack of them) from the 3+4+3 components of the three reuse sweeps."""
        shape = np.shape(c1[0])
        rho = np.zeros(shape + (4, 4), dtype=complex)
        rho[..., 0, 0], rho[..., 1, 1] = np.abs(c1[0]), np.abs(c1[2])
        rho[..., 2, 2], rho[..., 3, 3] = np.abs(c3[0]), np.abs(c3[2])
        upper = {(0, 1): c1[1], (0, 2): c2[0], (0, 3): c2[1], (1, 2): c2[2], (1, 3): c2[3], (2, 3): c3[1]}
        for (r, c), v in upper.items():
            rho[..., r, c] = v
            rho[..., c, r] = np.conj(v)
        return rho

    def calc_densitymatrix_reuse(self, plot_G2=None, return_counts=False, return_rho=False):
        """Three sweeps instead of ten (reference ``:301-373``).  ``plot_G2`` saves the ``t1``-resolved
        components as ``<plot_G2>.npy`` (plotting itself is out of scope)."""
        (t1, g1_t, g1), (_, g2_t, g2), (_, g3_t, g3) = self._reuse_sweeps()
        rho = self._assemble(g1, g2, g3)
        norm = np.trace(rho)
        if plot_G2 is not None:
            np.save("{}.npy".format(plot_G2), np.array([t1, *g1_t, *g2_t, *g3_t]))
        if return_rho:
            return concurrence(rho / norm), rho
        if return_counts:
            return concurrence(rho / norm), rho[0, 0], rho[1, 1], rho[2, 2], rho[3, 3], rho[0, 3]
        return concurrence(rho / norm)

    # ------------------------------------------------------------------ time-resolved entanglement
    def calc_timedep_data(self):
        """Full ``G2(t, tau)`` of the ten components, stacked (reference ``:375-389``)."""
        sweeps = self._reuse_sweeps(full=True)
        t1, t2 = sweeps[0][0], sweeps[0][1]
        return t1, t2, np.concatenate([s[4] for s in sweeps], axis=0)

    def integrate_g2_tau(self, t1, t2, G2_full):
        """``G2(tau) = int dt G2(t, tau)`` (reference ``:535-549``)."""
        return t2, np.trapezoid(G2_full, t1, axis=1)

    def integrate_timedep_G2(self, t1, t2, G2_full):
        """``G2(t) = int_0^t dt' int_0^{t-t'} dtau G2(t', tau)`` (reference ``:552-606``).  The
        reference's O(n_t^2 n_tau) Python loop becomes cumulative trapezoids plus one gather per
        ``t``; the sample sets are identical (all ``tau`` grid points ``<= t - t'``)."""
        n_c, n_t, n_tau = G2_full.shape
        seg = 0.5 * (G2_full[:, :, 1:] + G2_full[:, :, :-1]) * np.diff(t2)[None, None, :]
        cum = np.concatenate([np.zeros((n_c, n_t, 1), dtype=complex), np.cumsum(seg, axis=2)], axis=2)
        out = np.zeros((n_c, n_t), dtype=complex)
        rows = np.arange(n_t)
        for i in range(n_t):
            last = np.searchsorted(t2, t1[i] - t1[:i + 1], side="right") - 1   # last tau index <= t - t'
            inner = cum[:, rows[:i + 1], np.maximum(last, 0)]
            out[:, i] = np.trapezoid(inner, t1[:i + 1], axis=1)
        return t1, out

    def calc_timedependent_rho(self, plot_G2=None, t1=None, t2=None, G2_full=None, t=None, G2_t=None, add_norm=0,
                               mode="t", skip=0, return_G2=False):
        """Two-photon density matrix and concurrence resolved in detection time (``mode="t"``) or delay
        (``mode="tau"``), reference ``:391-437``."""
        if t is None or G2_t is None:
            if t1 is None or t2 is None or G2_full is None:
                t1, t2, G2_full = self.calc_timedep_data()
            if mode == "t":
                t, G2_t = self.integrate_timedep_G2(t1, t2, G2_full)
            if mode == "tau":
                t, G2_t = self.integrate_g2_tau(t1, t2, G2_full)
        t, G2_t = t[skip:], G2_t[:, skip:]
        rho_t = self._assemble(G2_t[0:3], G2_t[3:7], G2_t[7:10])
        rho_int = np.trapezoid(rho_t, t, axis=0)
        c_int = concurrence(rho_int / np.trace(rho_int).real)
        for k in range(4):      # uncorrelated background on the diagonal
            rho_t[:, k, k] += add_norm
        norm = np.trace(rho_t, axis1=1, axis2=2).real
        c_t = np.array([concurrence(rho_t[i] / norm[i]) for i in range(len(t))])
        if plot_G2 is not None:
            np.savez("{}.npz".format(plot_G2), t1=t1, t2=t2, G2_full=G2_full)
        if return_G2:
            return t, c_t, rho_t, norm, rho_int, c_int, G2_t
        return t, c_t, rho_t, norm, rho_int, c_int
