"""Two-photon time-bin entanglement of the biexciton-exciton cascade.

``TwoPhotonTimebinNew`` keeps the constructor, method names and return tuples of the reference's
``pyaceqd/timebin/twophoton_new.py:18-557,1015-1148``: the 4x4 two-photon density matrix in the
basis ``|ee>, |el>, |le>, |ll>`` is built from ten G2-type components.  The reference writes every
component out as its own thread-pool loop; here they are instances of two sweep shapes
(:mod:`pyaceqd_b200.sweeps`): a *tail sweep* over ``t1`` and a *triangular sweep* over pairs
``t1 <= t2`` of which only the final output value is used (``four_time`` ``:515-557``).  A whole
sweep -- for the triangular ones all ``n(n+1)/2`` trajectories -- is issued as GPU batches in
which trajectories sharing a prefix are forked from one trunk.

The time-local dynamical-map fast path (``calc_densitymatrix_tl`` ``:100-181``) is the mixin
:class:`pyaceqd_b200.timebin.twophoton_tl.TimeLocalTimebin` on top of the chain kernel.
"""
from __future__ import annotations

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.sweeps import at_time, run_sweep, run_sweep_arrays, tail_series
from pyaceqd_b200.timebin.timebin import TimeBin
from pyaceqd_b200.timebin.twophoton_tl import TimeLocalTimebin
from pyaceqd_b200.tools import concurrence, construct_t, simple_t_gaussian

temp_dir = constants.temp_dir

options_example = {"verbose": False, "delta_xd": 4, "gamma_e": 1 / 65, "lindblad": True, "temp_dir": temp_dir,
                   "phonons": False, "pt_file": "tls_dark_3.0nm_4k_th10_tmem20.48_dt0.02.ptr"}

_TRI_CHUNK = 4096     # trajectories per GPU batch of a triangular sweep


def _L(op):
    return {"operator": op, "applyFrom": "_left", "applyBefore": "false"}


def _R(op):
    return {"operator": op, "applyFrom": "_right", "applyBefore": "false"}


class TwoPhotonTimebinNew(TimeLocalTimebin, TimeBin):
    def __init__(self, system, sigma_x, sigma_xdag, sigma_b, sigma_bdag, *pulses, dt=0.02, dim=5, tb=800,
                 dt_small=0.1, n_tbig=10, dt_exp=None, simple_exp=True, gaussian_t=None, verbose=False, workers=15,
                 simple_t=False, options={}) -> None:
        super().__init__(system, *pulses, dt=dt, tb=tb, simple_exp=simple_exp, gaussian_t=gaussian_t,
                         verbose=verbose, workers=workers, options=options)
        self.gamma_e = options["gamma_e"]
        self.dim = dim
        self.prepare_operators(sigma_x=sigma_x, sigma_xdag=sigma_xdag, sigma_b=sigma_b, sigma_bdag=sigma_bdag,
                               verbose=verbose)
        if self.gaussian_t is not None:
            self.t1 = simple_t_gaussian(0, self.gaussian_t, self.tb, dt_small, n_tbig * dt_small, *self.pulses,
                                        decimals=1, exp_part=self.simple_exp)
        if self.gaussian_t is None or simple_t:
            self.t1 = construct_t(0, self.tb, dt_small, n_tbig * dt_small, dt_exp, *self.pulses,
                                  simple_exp=self.simple_exp)

    def prepare_operators(self, sigma_x, sigma_xdag, sigma_b, sigma_bdag, verbose=False):
        """Exciton (``x``) and biexciton (``b``) photon operators and the number operators built from them."""
        self.sigma_x, self.sigma_xdag = sigma_x, sigma_xdag
        self.sigma_b, self.sigma_bdag = sigma_b, sigma_bdag
        self.x_op = "(" + sigma_xdag + " * " + sigma_x + ")"
        self.b_op = "(" + sigma_bdag + " * " + sigma_b + ")"
        if verbose:
            print("sigma_x: {}, sigma_xdag: {}, x_op: {}".format(self.sigma_x, self.sigma_xdag, self.x_op))
            print("sigma_b: {}, sigma_bdag: {}, b_op: {}".format(self.sigma_b, self.sigma_bdag, self.b_op))

    def calc_timedynamics(self, output_ops=None):
        opts = self.options.copy()
        if output_ops is not None:
            opts["output_ops"] = output_ops
        return self.system(0, 2 * self.tb, *self.pulses, **opts)

    # ------------------------------------------------------------------ sweep shapes
    def _n_tau(self):
        return int(self.tb / self.dt)

    def _tail_sweep(self, output_ops, mtos_with_shift, tend, tail_of):
        """One job per ``t1``; ``mtos_with_shift`` = [(mto, shift)] places each operator at ``t1 + shift``."""
        jobs = [{"tend": tend, "mtos": [at_time(m, t + sh) for m, sh in mtos_with_shift],
                 "output_ops": output_ops, "tail": tail_of(t)} for t in self.t1]
        return run_sweep(self.system, jobs, options=self.options, workers=self.workers)

    def _triangular_sweep(self, output_ops, mtos_with_when, tend_when, special_first):
        """All pairs ``t1[i] <= t2 = t1[i + j]``: operators at ``t1``, ``t2`` or ``t1 + tb`` in the given
        (file) order -- the order resolves coinciding times (``:436-438``) -- and the trajectory ends at
        ``t2 + tb`` or ``t1 + tb``.  Returns the matrix of final output values: the first output, or
        the second one on the diagonal ``t2 = t1`` if ``special_first`` (``:549-552``)."""
        t1 = np.asarray(self.t1, dtype=float)
        n = len(t1)
        when = {"t1": lambda a, b: a, "t2": lambda a, b: b, "t1+tb": lambda a, b: a + self.tb,
                "t2+tb": lambda a, b: b + self.tb}
        vals = np.zeros((n, n), dtype=complex)
        ii, jj = np.triu_indices(n)                    # row-major: all t2 of one t1 are neighbours
        templates = [m for m, _ in mtos_with_when]
        # whole t1 rows per GPU batch: the pairs of one t1 share their stretch up to t2 (two-level forking)
        row_end = np.cumsum(n - np.arange(n))
        lo = 0
        while lo < len(ii):
            hi = int(row_end[min(np.searchsorted(row_end, lo + _TRI_CHUNK), n - 1)])
            a, b = t1[ii[lo:hi]], t1[jj[lo:hi]]
            times = np.stack([when[w](a, b) for _, w in mtos_with_when], axis=1)
            res = run_sweep_arrays(self.system, 0, when[tend_when](a, b), templates, times, output_ops=output_ops,
                                   tails=1, options=self.options, workers=self.workers)
            last = res.last_rows() if hasattr(res, "last_rows") else np.array([[x[-1] for x in r[1:]] for r in res])
            diag = ii[lo:hi] == jj[lo:hi]
            vals[ii[lo:hi], jj[lo:hi]] = np.where(diag, last[:, 1], last[:, 0]) if special_first else last[:, 0]
            lo = hi
        return vals

    def _integrate_triangular(self, vals):
        t1 = self.t1
        g2 = np.array([np.trapezoid(vals[i, i:], t1[i:]) for i in range(len(t1))])
        return g2, np.trapezoid(g2, t1) * self.gamma_e ** 2

    def four_time(self, output_ops, sigma_1, sigma_2, sigma_3):
        """Operators at ``t1``, ``t2`` and ``t1 + tb``; end at ``t2 + tb`` (reference ``:515-557``)."""
        vals = self._triangular_sweep(output_ops, [(sigma_1, "t1"), (sigma_2, "t2"), (sigma_3, "t1+tb")],
                                      "t2+tb", special_first=True)
        g2, total = self._integrate_triangular(vals)
        n = len(self.t1)
        aligned = np.zeros((n, n), dtype=complex)          # reference stores row i right-aligned
        for i in range(n):
            aligned[i, i:] = vals[i, i:]
        return self.t1, g2, total, aligned

    # ------------------------------------------------------------------ diagonal of the density matrix
    def rho_ee_ee(self, add_time=0, use_second_zero=False):
        """Both photons in the same bin (shifted by ``add_time``): biexciton photon first, then -- to
        cover re-excitation -- exciton photon first (reference ``:201-278``)."""
        t1, n_tau = self.t1, self._n_tau()
        t2 = np.linspace(0, self.tb, n_tau + 1)

        def part(output_ops, first, first_dag):
            res = self._tail_sweep(output_ops, [(_L(first), add_time), (_R(first_dag), add_time)],
                                   self.tb + add_time, lambda t: n_tau - int(t / self.dt) + 1)
            g2 = np.zeros(len(t1))
            grid = np.zeros((len(t1), len(t2)))
            for i, (t, r) in enumerate(zip(t1, res)):
                col = np.abs(tail_series(r, n_tau - int(t / self.dt)))
                g2[i] = np.trapezoid(col, t2[:len(col)])
                grid[i, -len(col):] = col
            return g2, grid

        x_num = self.sigma_xdag + "*" + self.sigma_x
        g2_1, grid_1 = part([x_num, self.sigma_bdag + "*" + x_num + "*" + self.sigma_b], self.sigma_b, self.sigma_bdag)
        if use_second_zero:
            return t1, t2, g2_1, np.trapezoid(g2_1, t1) * self.gamma_e ** 2, g2_1, g2_1 * 0, grid_1
        # t2 = t1 is already covered by the first ordering -> a vanishing tau=0 operator for the second
        g2_2, grid_2 = part([self.sigma_bdag + "*" + self.sigma_b, "0*" + self.sigma_xdag], self.sigma_x, self.sigma_xdag)
        g2 = g2_1 + g2_2
        return t1, t2, g2, np.trapezoid(g2, t1) * self.gamma_e ** 2, g2_1, g2_2, grid_1 + grid_2

    def rho_ll_ll(self, use_second_zero=False):
        return self.rho_ee_ee(add_time=self.tb, use_second_zero=use_second_zero)

    def rho_el_el(self, output_ops=None, sigma_X=None, sigma_Xdag=None):
        """First photon in the early bin at ``t1``, second anywhere in the late bin (reference
        ``:286-348``); the bins touch only at ``t1 = tb``, where the tau=0 operator is used."""
        if output_ops is None:
            x_num = self.sigma_xdag + "*" + self.sigma_x
            output_ops = [x_num, self.sigma_bdag + "*" + x_num + "*" + self.sigma_b]
        sigma_X = _L(self.sigma_b) if sigma_X is None else sigma_X
        sigma_Xdag = _R(self.sigma_bdag) if sigma_Xdag is None else sigma_Xdag
        t1, n_tau = self.t1, self._n_tau()
        t2 = np.linspace(0, self.tb, n_tau + 1)
        res = self._tail_sweep(output_ops, [(sigma_X, 0), (sigma_Xdag, 0)], 2 * self.tb, lambda t: n_tau + 1)
        g2 = np.zeros(len(t1))
        for i, r in enumerate(res):
            col = np.abs(r[1][-n_tau - 1:])
            if i == len(t1) - 1:
                col[0] = np.abs(r[2][-n_tau - 1])
            g2[i] = np.trapezoid(col, t2)
        return t1, g2, np.trapezoid(g2, t1) * self.gamma_e ** 2

    def rho_le_le(self):
        b_num = self.sigma_bdag + "*" + self.sigma_b
        return self.rho_el_el(output_ops=[b_num, self.sigma_xdag + "*" + b_num + "*" + self.sigma_x],
                              sigma_X=_L(self.sigma_x), sigma_Xdag=_R(self.sigma_xdag))

    # ------------------------------------------------------------------ coherences
    def rho_ee_ll(self, use_second_zero=False):
        """``<ee|rho|ll>``: four times on two axes, the late pair shifted by exactly ``tb`` (reference
        ``:368-393``; assumes identical pulses in both bins)."""
        t1, g2_1, v1, grid_1 = self.four_time([self.sigma_x, self.sigma_x + "*" + self.sigma_b],
                                              _R(self.sigma_bdag), _R(self.sigma_xdag), _L(self.sigma_b))
        if use_second_zero:
            return t1, g2_1, v1, g2_1, g2_1 * 0, grid_1
        t1, g2_2, v2, grid_2 = self.four_time([self.sigma_bdag, self.sigma_b + "*" + self.sigma_x],
                                              _R(self.sigma_xdag), _R(self.sigma_bdag), _L(self.sigma_x))
        return t1, g2_1 + g2_2, v1 + v2, g2_1, g2_2, grid_1 + grid_2

    def rho_ee_el(self, operators=None):
        """``<ee|rho|el>`` (reference ``:395-506``).  ``operators = [out, left, right_1, right_2]``."""
        out, left, right_1, right_2 = operators if operators is not None else \
            (self.sigma_x, self.sigma_b, self.sigma_bdag, self.sigma_xdag)
        if operators is not None and len(operators) != 4:
            raise ValueError("operators must be a list of length 4")
        # t1 <= t2: pair (left, right_1) at t1, right_2 at t2, end t2 + tb
        v1 = self._triangular_sweep([out], [(_L(left), "t1"), (_R(right_1), "t1"), (_R(right_2), "t2")],
                                    "t2+tb", special_first=False)
        # t2 <= t1: right_2 first at t1, the pair at t2, end t1 + tb
        v2 = self._triangular_sweep([out], [(_R(right_2), "t1"), (_L(left), "t2"), (_R(right_1), "t2")],
                                    "t1+tb", special_first=False)
        (g1, tot1), (g2, tot2) = self._integrate_triangular(v1), self._integrate_triangular(v2)
        return self.t1, g1 + g2, tot1 + tot2, g1, g2

    def rho_ee_le(self):
        return self.rho_ee_el(operators=[self.sigma_b, self.sigma_x, self.sigma_xdag, self.sigma_bdag])

    def rho_el_le(self):
        """``<el|rho|le>`` from two ``four_time`` sweeps (reference ``:1015-1029``)."""
        t1, g1, tot1, _ = self.four_time([self.sigma_xdag, self.sigma_xdag + "*" + self.sigma_b],
                                         _R(self.sigma_bdag), _L(self.sigma_x), _L(self.sigma_b))
        t1, g2, tot2, _ = self.four_time([self.sigma_b, self.sigma_xdag + "*" + self.sigma_b],
                                         _L(self.sigma_x), _R(self.sigma_bdag), _R(self.sigma_xdag))
        return t1, g1 + g2, tot1 + tot2, g1, g2

    def rho_el_ll(self, calc_lell=False):
        """``<el|rho|ll>`` (``<le|rho|ll>`` with exciton and biexciton operators exchanged), reference
        ``:1031-1148``: a tail sweep for ``t1 <= t2`` and a triangular sweep for ``t2 <= t1``."""
        x, xd, b, bd = self.sigma_x, self.sigma_xdag, self.sigma_b, self.sigma_bdag
        if calc_lell:
            x, xd, b, bd = b, bd, x, xd
        t1, n_tau = self.t1, self._n_tau()
        t2 = np.linspace(0, self.tb, n_tau + 1)
        num = xd + "*" + x
        res = self._tail_sweep([num, num + "*" + b], [(_R(bd), 0), (_L(b), self.tb)], 2 * self.tb,
                               lambda t: n_tau - int(t / self.dt) + 1)
        g1 = np.zeros(len(t1), dtype=complex)
        for i, (t, r) in enumerate(zip(t1, res)):
            col = tail_series(r, n_tau - int(t / self.dt))
            g1[i] = np.trapezoid(col, t2[:len(col)])
        tot1 = np.trapezoid(g1, t1) * self.gamma_e ** 2
        vals = self._triangular_sweep([b, xd + "*" + b + "*" + x],
                                      [(_R(bd), "t2"), (_L(x), "t1+tb"), (_R(xd), "t1+tb")], "t2+tb",
                                      special_first=True)
        g2, tot2 = self._integrate_triangular(vals)
        return t1, g1 + g2, tot1 + tot2, g1, g2

    def rho_le_ll(self):
        return self.rho_el_ll(calc_lell=True)

    # ------------------------------------------------------------------ assembly
    def calc_densitymatrix(self, save_dm=False, save_all=False, filename="densitymatrix", verbose=False,
                           reduced=False, use_second_zero=False):
        """Returns ``(concurrence, unnormalised density matrix)`` (reference ``:38-98``); ``reduced``
        keeps only the populations and ``<ee|rho|ll>``."""
        rho = np.zeros((4, 4), dtype=complex)
        t, _, c_eeee, rho[0, 0], c_eeee_1, c_eeee_2, _ = self.rho_ee_ee(use_second_zero=use_second_zero)
        _, c_elel, rho[1, 1] = self.rho_el_el()
        _, c_lele, rho[2, 2] = self.rho_le_le()
        _, _, c_llll, rho[3, 3], c_llll_1, c_llll_2, _ = self.rho_ll_ll(use_second_zero=use_second_zero)
        _, c_eell, rho[0, 3], c_eell_1, c_eell_2, _ = self.rho_ee_ll(use_second_zero=use_second_zero)
        zero = 0 * c_eeee
        parts = {k: (zero, zero, zero) for k in ("eeel", "eele", "elle", "elll", "lell")}
        if not reduced:
            for key, (r, c), fn in (("eeel", (0, 1), self.rho_ee_el), ("eele", (0, 2), self.rho_ee_le),
                                    ("elle", (1, 2), self.rho_el_le), ("elll", (1, 3), self.rho_el_ll),
                                    ("lell", (2, 3), self.rho_le_ll)):
                _, g, rho[r, c], g_1, g_2 = fn()
                parts[key] = (g, g_1, g_2)
        for r in range(4):
            for c in range(r + 1, 4):
                rho[c, r] = np.conj(rho[r, c])
        norm = np.trace(rho)
        if save_dm or save_all:
            np.save(filename + "_dm.npy", rho)
        if save_all:
            np.save(filename + "_t.npy", t)
            order = ("eeel", "eele", None, "elle", "elll", "lell")
            full = [c_eeee, c_elel, c_lele, c_llll] + [c_eell if k is None else parts[k][0] for k in order]
            np.save(filename + "_components.npy", np.stack(full, axis=0))
            for which, first in ((1, [c_eeee_1, c_llll_1]), (2, [c_eeee_2, c_llll_2])):
                sub = first + [(c_eell_1 if which == 1 else c_eell_2) if k is None else parts[k][which]
                               for k in order]
                np.save(filename + "_components_{}.npy".format(which), np.stack(sub, axis=0))
        if verbose:
            fmt = {'complex_kind': lambda z: "%.3f+%.3fj" % (z.real, z.imag)}
            print("density matrix:\n" + np.array2string(rho, formatter=fmt))
            print("normalized density matrix:\n" + np.array2string(rho / norm, formatter=fmt))
        return concurrence(rho / norm), rho
