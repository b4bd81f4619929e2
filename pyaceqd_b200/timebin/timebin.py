"""Common base of the time-bin entanglement workflows: holds the system adapter, the bin width
``tb`` and the drive sampled once for both bins (reference ``pyaceqd/timebin/timebin.py:7-98``)."""
from __future__ import annotations

import os

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.tools import export_csv

temp_dir = constants.temp_dir


def _sample_xy(pulses, grid):
    fx = np.zeros_like(grid, dtype=complex)
    fy = np.zeros_like(grid, dtype=complex)
    for p in pulses:
        f = p.get_total(grid)
        fx, fy = fx + p.polar_x * f, fy + p.polar_y * f
    return fx, fy


class TimeBin():
    def __init__(self, system, *pulses, dt=0.02, tb=800, simple_exp=True, gaussian_t=None, verbose=False,
                 workers=15, t_simul=None, options={}) -> None:
        self.system = system
        self.dt = dt
        self.options = dict(options)
        self.options["dt"] = dt
        self.tb = tb
        self.simple_exp = simple_exp
        self.gaussian_t = gaussian_t
        self.pulses = pulses
        self.workers = workers
        self._own_files = []
        if "temp_dir" not in options:
            print("temp_dir not included in options, setting to temp_dir specified in constants")
            self.options["temp_dir"] = temp_dir
        self.temp_dir = self.options["temp_dir"]
        if self.options.get("pulse_file_x") is None or self.options.get("pulse_file_y") is None:
            self.prepare_pulsefile(verbose=verbose, t_simul=t_simul)
            self.options["pulse_file_x"] = self.pulse_file_x
            self.options["pulse_file_y"] = self.pulse_file_y
        else:
            self.pulse_file_x = self.options["pulse_file_x"]
            self.pulse_file_y = self.options["pulse_file_y"]

    def _write(self, name, grid, f, verbose):
        path = self.temp_dir + name.format(id(self))     # object id: concurrent instances must not collide
        export_csv(path, grid, f.real, f.imag, precision=8, delimit=' ', verbose=verbose)
        self._own_files.append(path)
        return path

    def prepare_pulsefile(self, verbose=False, t_simul=None):
        """Both bins in one file pair: ``[0, 2.1 tb)`` (or ``t_simul``) on ``dt/5`` (reference ``:32-47``)."""
        grid = np.arange(0, 2.1 * self.tb if t_simul is None else t_simul, step=self.dt / 5)
        fx, fy = _sample_xy(self.pulses, grid)
        self.pulse_file_x = self._write("timebin_pulse_x_{}.dat", grid, fx, verbose)
        self.pulse_file_y = self._write("timebin_pulse_y_{}.dat", grid, fy, verbose)

    def prepare_puslefile_tls(self, verbose=False):
        """Separate file pairs for the early and the late bin, the late one shifted to start at 0 so
        that a propagation from t=0 sees the right carrier phase (reference ``:49-86``; the
        misspelled name is the reference's)."""
        for k, (lo, hi) in enumerate(((0, self.tb), (self.tb, 2 * self.tb)), start=1):
            grid = np.arange(lo, hi, step=self.dt / 5)
            mine = [p for p in self.pulses if (p.t0 < self.tb) == (k == 1)]
            fx, fy = _sample_xy(mine, grid)
            setattr(self, "pulse_file_x{}".format(k), self._write("timebin_pulse_x_tb%d_{}.dat" % k, grid - lo, fx, verbose))
            setattr(self, "pulse_file_y{}".format(k), self._write("timebin_pulse_y_tb%d_{}.dat" % k, grid - lo, fy, verbose))

    def __del__(self):
        for f in getattr(self, "_own_files", []):
            if os.path.exists(f):
                os.remove(f)
