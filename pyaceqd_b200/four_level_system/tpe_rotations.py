"""Two-photon-excitation area sweeps of the biexciton system: exciton / biexciton occupations (or
photon counts) versus pulse area.  ``TPERotations`` follows the reference's
``pyaceqd/four_level_system/tpe_rotations.py:18-243``; the sweep machinery is shared with
:class:`pyaceqd_b200.two_level_system.rabi_rotations.RabiRotations`."""
from __future__ import annotations

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.two_level_system.rabi_rotations import _AreaSweep

temp_dir = constants.temp_dir


class TPERotations(_AreaSweep):
    prefix = "tpe_"
    n_columns = 3
    integrate_window_factor = 10
    integrate_extra = 100.0

    def __init__(self, dt=0.1, tau=5, delta_xy=0, delta_b=4, area_max=30, n_area=150, gamma_e=1 / 100, phonons=False,
                 temperature=4, ae=5, ah_ratio=1.15, J_from_file=None, phonon_factor=1, t_mem=6.1) -> None:
        super().__init__(dt, tau, area_max, n_area, gamma_e, phonons, temperature, ae, ah_ratio, J_from_file,
                         phonon_factor, t_mem, temp_dir)
        self.delta_xy, self.delta_b = delta_xy, delta_b
        self.options.update({"delta_xy": delta_xy, "delta_b": delta_b})

    def _system(self, *a, **kw):
        return biexciton(*a, **kw)

    def _generate_kwargs(self):
        return {"delta_xy": self.delta_xy, "delta_b": self.delta_b}

    def _pulse_file_kw(self):
        return "pulse_file_x"

    def _read(self, res, integrate):
        t, g, x, y, b = res
        if not integrate:
            return np.array([x[-1].real, y[-1].real, b[-1].real])
        t = np.real(t)
        # the biexciton emits two photons
        return self.gamma_e * np.array([np.trapezoid(np.real(x), t), np.trapezoid(np.real(y), t),
                                        2 * np.trapezoid(np.real(b), t)])
