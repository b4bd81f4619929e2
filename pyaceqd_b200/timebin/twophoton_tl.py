"""Time-local-map route of the two-photon time-bin density matrix (mixin of ``TwoPhotonTimebinNew``).

Reference: ``pyaceqd/timebin/twophoton_new.py`` -- ``calc_densitymatrix_tl`` ``:100-181``,
``_calc_dynmaps`` ``:559-597``, ``_calc_binary_steps`` ``:599-613``, ``eell_tl_f`` ``:629-670``,
``eell_tl_8ops`` ``:672-704``, ``eightops_fortran`` ``:706-717``, ``get_initial_state`` ``:719-728``.

Two dynamical-map runs (early and late bin, ``gaussian_t + 10`` ps each) give the explicit
time-local maps around the pulses; beyond them the last map is stationary and is fast-forwarded by
binary powers.  Every ``(t1, t2)`` element of a G2 component is then a chain of matrix-vector
products with operator insertions -- the reference's Fortran ``four_time_8op``; here one launch of
the chain kernel for all pairs (``pyaceqd_b200/timebin/timebin_tl.py``).  As in the reference the
maps are handed over transposed-to-``[NL, NL, n]`` and conjugated, which makes the column-major
reading of the row-major vectorised density matrix consistent for Hermitian states.  Not valid
with phonons (the reference prints the same warning).
"""
from __future__ import annotations

import numpy as np

from pyaceqd_b200.timebin import timebin_tl
from pyaceqd_b200.tools import calc_tl_dynmap_pseudo, concurrence, op_to_matrix


class TimeLocalTimebin:
    """Needs ``system, options, dt, tb, gaussian_t, t1, dim, gamma_e, sigma_*`` of the host class."""

    def get_initial_state(self):
        init = "|0><0|_{dim}".format(dim=self.dim)
        if "initial" in self.options:
            init = self.options["initial"]
            print("Using initial state from options:", init)
        else:
            print("Warning: no initial state given, assuming ground state.")
        return op_to_matrix(init)

    def _calc_binary_steps(self, tl_map):
        """``E, E^2, E^4, ...`` up to one bin length (reference ``:599-613``)."""
        n_bin = int(np.log2(int(self.tb / self.dt))) + 1
        out = np.zeros((n_bin,) + tl_map.shape, dtype=complex)
        out[0] = tl_map
        for i in range(1, n_bin):
            out[i] = out[i - 1] @ out[i - 1]
        return out

    def _calc_dynmaps(self):
        if self.options.get("phonons"):
            print("Phonons are enabled in the options. Correlation functions will give wrong results.")
        self.prepare_puslefile_tls()
        tls = []
        for fx, fy in ((self.pulse_file_x1, self.pulse_file_y1), (self.pulse_file_x2, self.pulse_file_y2)):
            opts = dict(self.options)
            opts["pulse_file_x"], opts["pulse_file_y"] = fx, fy
            result, dm = self.system(0, self.gaussian_t + 10, calc_dynmap=True, **opts)
            tls.append(calc_tl_dynmap_pseudo(dm, np.round(np.real(result[0]), 6)))
        if len(tls[0]) != len(tls[1]):
            print("Warning: time axes of dyn. maps are not the same length. Check if anything is wrong.")
        self.dm_tl1, self.dm_tl2 = tls
        tl_map = tls[0][-1]
        self.precalc_tls = self._calc_binary_steps(tl_map)
        return tl_map, self.dm_tl1, self.dm_tl2

    @staticmethod
    def _f(maps):
        """``[n, NL, NL]`` row-major maps -> the ``[NL, NL, n]`` conjugated layout the chain routines take."""
        return np.asfortranarray(np.asarray(maps).transpose(1, 2, 0).conjugate())

    def _integrate_pairs(self, grid):
        g2 = np.array([np.trapezoid(grid[i, i:], self.t1[i:]) for i in range(len(self.t1))])
        return g2, np.trapezoid(g2, np.round(self.t1, 6)) * self.gamma_e ** 2

    def eightops_fortran(self, rho0, operators, precalc_tls, dm_1, dm_2, early_only=False, late_t1_only=False):
        """One G2 component on the ``t1 <= t2`` triangle from eight (left, right) operators at the two early
        and the two late times (reference ``:706-717``)."""
        dim = rho0.shape[0]
        grid = timebin_tl.four_time_8op(dm_1, dm_2, rho0.reshape(dim * dim), np.round(self.t1, 6), precalc_tls,
                                        np.round(self.dt, 6), dim, *operators, early_only, late_t1_only, self.tb)
        g2, total = self._integrate_pairs(grid)
        return np.round(self.t1, 6), g2, total, grid

    def eell_tl_f(self):
        """``<ee|rho|ll>`` with the four-operator routine (reference ``:629-670``)."""
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        dim = rho0.shape[0]
        ops = [op_to_matrix(o) for o in (self.sigma_bdag, self.sigma_xdag, self.sigma_b, self.sigma_x)]
        grid = timebin_tl.four_time(self._f(dm_1), self._f(dm_2), rho0.reshape(dim * dim), np.round(self.t1, 6),
                                    self._f(self.precalc_tls), np.round(self.dt, 6), dim, *ops, self.tb)
        g2, total = self._integrate_pairs(grid)
        return np.round(self.t1, 6), g2, total, grid

    # ------------------------------------------------------------------ the reference's Python-level routines
    def fast_propagate(self, rho, n):
        """``E^n rho`` from the binary powers of the stationary map (reference ``:730-735``)."""
        for i, bit in enumerate(reversed(np.binary_repr(int(n)))):
            if bit == "1":
                rho = self.precalc_tls[i] @ rho
        return rho

    def propagate_tb_new(self, t_start, t_stop, rho, dm_tl, verbose=False):
        """Explicit maps ``dm_tl[n]`` while they last, then the stationary fast-forward (reference ``:737-759``);
        ``rho`` is the row-major vectorised density matrix, maps are ``[n, NL, NL]``."""
        n_start = int(np.round(np.round(t_start, 6) / self.dt))
        n_steps = int(np.round(np.round(t_stop, 6) / self.dt)) - n_start
        steps_dm = max(0, min(len(dm_tl) - n_start, n_steps))
        if verbose:
            print(f"propagate from {t_start} to {t_stop} using {n_steps} steps of {self.dt}, of which {steps_dm} are from dm")
        for k in range(steps_dm):
            rho = dm_tl[n_start + k] @ rho
        return self.fast_propagate(rho, max(0, n_steps - steps_dm))

    def four_time_tl(self, sigma_1, sigma_2, sigma_3, sigma_4, supply_mats=False):
        """``Tr[s4 s3 (... rho s1 ... s2 ...)]`` with ``s1`` at ``t1`` and ``s2`` at ``t2`` from the right in the
        early bin, ``s3`` at ``t1 + tb`` and ``s4`` at ``t2 + tb`` from the left in the late bin, for all
        ``t1 <= t2`` (reference ``:925-1013``, a double Python loop there; here the chain kernel through the
        four-operator routine).  Returns ``(t1, G2(t1), integral * gamma_e^2, G2(t1, t2))``."""
        ops = [np.asarray(o, dtype=complex) if supply_mats else op_to_matrix(o) for o in (sigma_1, sigma_2, sigma_3, sigma_4)]
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        dim = rho0.shape[0]
        self.t1 = np.round(self.t1, 6)
        grid = timebin_tl.four_time(self._f(dm_1), self._f(dm_2), rho0.reshape(dim * dim), self.t1,
                                    self._f(self.precalc_tls), np.round(self.dt, 6), dim, *ops, self.tb)
        g2, total = self._integrate_pairs(grid)
        return self.t1, g2, total, grid

    def eell_tl(self):
        """``<ee|rho|ll>`` through :meth:`four_time_tl` (reference ``:615-627``)."""
        t1, g2, eell, grid = self.four_time_tl(self.sigma_bdag, self.sigma_xdag, self.sigma_b, self.sigma_x)
        return t1, g2, eell, g2, g2 * 0, grid

    def dynamics_tl(self):
        """Density matrix on the fine grid through both bins from the time-local maps (reference ``:761-790``)."""
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        dim = rho0.shape[0]
        t = np.arange(0, 2 * self.tb, self.dt)
        rho_t = np.zeros((len(t), dim, dim), dtype=complex)
        rho_t[0] = rho0
        n_tb = int(self.tb / self.dt)
        for i in range(len(t) - 1):
            k, dm = (i, dm_1) if i < n_tb else (i - n_tb, dm_2)
            rho_t[i + 1] = self.propagate_tb_new(k * self.dt, (k + 1) * self.dt, rho_t[i].reshape(dim * dim), dm).reshape(dim, dim)
        return t, rho_t

    def dynamics_tl_t1(self):
        """Density matrix on the ``t1`` grid through both bins (reference ``:822-843``)."""
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        dim = rho0.shape[0]
        t1 = np.round(self.t1, 6)
        # real maps of a Hermiticity-preserving evolution: the unconjugated [NL, NL, n] layout, as the reference passes it
        res = timebin_tl.dynamics_t1(np.asarray(dm_1).transpose(1, 2, 0), np.asarray(dm_2).transpose(1, 2, 0),
                                     rho0.reshape(dim * dim), t1, np.asarray(self.precalc_tls).transpose(1, 2, 0),
                                     self.dt, dim, self.tb)
        return np.concatenate((t1, t1[1:] + self.tb)), np.array([res[:, k].reshape(dim, dim) for k in range(res.shape[1])])

    def dynamics_tl_t1_t2_f(self, _t1, _t2, sigma_1, sigma_2, sigma_3, take_IDs=False):
        """As :meth:`dynamics_tl_t1` with ``sigma_1`` / ``sigma_2`` from the right at ``_t1`` / ``_t2`` in the early bin
        and ``sigma_3`` from the left at ``_t1 + tb`` (reference ``:890-922``)."""
        rho0 = self.get_initial_state()
        dim = rho0.shape[0]
        ops = [np.eye(dim, dtype=complex)] * 3 if take_IDs else [op_to_matrix(o) for o in (sigma_1, sigma_2, sigma_3)]
        if not hasattr(self, "dm_tl1"):
            self._calc_dynmaps()
        t1 = np.round(self.t1, 6)
        res = timebin_tl.dynamics_t1_t2(self._f(self.dm_tl1), self._f(self.dm_tl2), _t1, _t2, rho0.reshape(dim * dim), t1,
                                        self._f(self.precalc_tls), self.dt, dim, self.tb, *ops)
        out = np.array([res[:, k].reshape(dim, dim).T for k in range(res.shape[1])])
        return np.concatenate((t1, t1[1:] + self.tb)), out

    def dynamics_tl_t1_t2(self, t1, t2, sigma_1, sigma_2, sigma_3, take_IDs=False):
        """The Python-level twin of :meth:`dynamics_tl_t1_t2_f` (reference ``:845-888``) on the ``t1`` grid of this object."""
        return self.dynamics_tl_t1_t2_f(t1, t2, sigma_1, sigma_2, sigma_3, take_IDs=take_IDs)

    def eell_tl_8ops(self):
        """The same element through the eight-operator routine (reference ``:672-704``)."""
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        I = np.eye(rho0.shape[0])
        ops = [I, op_to_matrix(self.sigma_bdag), I, op_to_matrix(self.sigma_xdag),
               op_to_matrix(self.sigma_b), I, op_to_matrix(self.sigma_x), I]
        return self.eightops_fortran(rho0, ops, self._f(self.precalc_tls), self._f(dm_1), self._f(dm_2))

    def calc_densitymatrix_tl(self, save_dm=False, filename="densitymatrix_tl", verbose=False, reduced=True):
        """Two-photon density matrix from time-local maps (reference ``:100-181``): without the
        second time ordering ``t2 <= t1``; ``reduced`` keeps the populations and ``<ee|rho|ll>``.
        Returns ``(concurrence, density matrix, normalised density matrix)``."""
        tl_map, dm_1, dm_2 = self._calc_dynmaps()
        rho0 = self.get_initial_state()
        pre, d1, d2 = self._f(self.precalc_tls), self._f(dm_1), self._f(dm_2)
        x, xd = op_to_matrix(self.sigma_x), op_to_matrix(self.sigma_xdag)
        b, bd = op_to_matrix(self.sigma_b), op_to_matrix(self.sigma_bdag)
        I = np.eye(rho0.shape[0])
        # (early t1 left, right, early t2 left, right, late t1 left, right, late t2 left, right), flags
        comps = {(0, 0): ([b, bd, x, xd, I, I, I, I], dict(early_only=True)),
                 (1, 1): ([b, bd, I, I, I, I, x, xd], {}),
                 (2, 2): ([x, xd, I, I, I, I, b, bd], {}),
                 (3, 3): ([I, I, I, I, b, bd, x, xd], {}),
                 (0, 3): ([I, bd, I, xd, b, I, x, I], {})}
        if not reduced:
            comps.update({(0, 1): ([b, bd, I, xd, I, I, I, x], {}),
                          (0, 2): ([I, bd, x, xd, I, b, I, I], dict(late_t1_only=True)),
                          (1, 2): ([I, bd, x, I, xd, I, I, b], {}),
                          (1, 3): ([I, bd, I, I, b, I, x, xd], {}),
                          (2, 3): ([I, I, I, xd, b, bd, I, x], {})})
        rho = np.zeros((4, 4), dtype=complex)
        for (r, c), (ops, flags) in comps.items():
            _, _, val, _ = self.eightops_fortran(rho0=rho0, operators=ops, precalc_tls=pre, dm_1=d1, dm_2=d2, **flags)
            rho[r, c] = val.real if r == c else val
            if r != c:
                rho[c, r] = np.conj(val)
        norm = np.trace(rho)
        if save_dm:
            np.save(filename + "_dm.npy", rho)
        return concurrence(rho / norm), rho, rho / norm
