"""Pulse-area sweeps ("Rabi rotations"): final or time-integrated exciton occupation versus pulse area.

``RabiRotations`` follows the reference's ``pyaceqd/two_level_system/rabi_rotations.py:16-228``
(constructor, ``generate_pt``, ``calc_timedynamics``, ``get_rabi_rotations``, CSV caching of the
result).  The reference submits one ACE run per area to a thread pool (``:172-198``); here the
whole sweep is one GPU batch.  ``_AreaSweep`` is shared with
:class:`pyaceqd_b200.four_level_system.tpe_rotations.TPERotations`.  Plotting and the FFT pulse
carving runs through :mod:`pyaceqd_b200.pulsegenerator` (fields handed over in memory, no pulse files on disk).
"""
from __future__ import annotations

import os

import numpy as np

import pyaceqd_b200.constants as constants
from pyaceqd_b200.batch import BatchExecutor, wait
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.tools import export_csv
from pyaceqd_b200.two_level_system.tls import tls

hbar = constants.hbar
temp_dir = constants.temp_dir


class _AreaSweep():
    """One trajectory per pulse area; subclasses name the adapter and what is read from its outputs."""
    prefix = "rabi_"
    n_columns = 1                 # result rows (e.g. x | x, y, b)
    integrate_window_factor = 11  # integrate mode propagates to round(factor / gamma_e)
    integrate_extra = 0.0

    def __init__(self, dt, tau, area_max, n_area, gamma_e, phonons, temperature, ae, ah_ratio, J_from_file,
                 phonon_factor, t_mem, temp_dir):
        self.dt, self.tau = dt, tau
        self.areas = np.linspace(0, area_max, n_area)
        self.gamma_e, self.phonons, self.temperature = gamma_e, phonons, temperature
        self.ae, self.ah_ratio, self.J_from_file = ae, ah_ratio, J_from_file
        self.phonon_factor, self.t_mem = phonon_factor, t_mem
        if J_from_file is not None:
            self.pt_name = J_from_file.split(".")[0] + ".ptr"
        else:
            self.pt_name = "pt_T{:.1f}K_AE{:.1f}_AHratio{:.2f}_coupl{:.1f}_dt{:.2f}_tmem{:.1f}.ptr".format(
                temperature, ae, ah_ratio, phonon_factor, dt, t_mem)
        # the reference lists ACE's four files; this engine stores one container under pt_name
        self.full_names = [self.pt_name] + [self.pt_name + s for s in ("_initial", "_initial_0", "_repeated", "_repeated_0")]
        self.options = {"gamma_e": gamma_e, "dt": dt, "phonons": phonons, "temp_dir": temp_dir, "pt_file": self.pt_name}
        if os.path.exists(self.pt_name):
            print("Warning: pt files already exist")

    # ---- to be provided by subclasses
    def _system(self, *a, **kw):
        raise NotImplementedError

    def _generate_kwargs(self):
        return {}

    def _pulse_file_kw(self):
        return "pulse_file"

    def _read(self, res, integrate):
        raise NotImplementedError

    # ---- shared
    def delete_pt_files(self):
        for name in self.full_names:
            if os.path.exists(name):
                os.remove(name)

    def generate_pt(self):
        """Build and cache the process tensor for the configured bath (reference ``:67-78``)."""
        p = ChirpedPulse(tau_0=self.tau, e_start=0, alpha=0, e0=1, polar_x=1.0, t0=4 * self.tau)
        self._system(0, 8 * self.tau, p, dt=self.dt, t_mem=self.t_mem, lindblad=False, phonons=True, ae=self.ae,
                     temperature=self.temperature, prepare_only=False, pt_file=self.pt_name,
                     **self._generate_kwargs())

    def _ensure_pt(self):
        if self.phonons and not os.path.exists(self.pt_name):
            self.generate_pt()

    def calc_timedynamics(self, tau, area, path="", save=False, plot_pulse=False, detuning=0, tend=None, plot=False,
                          plotlims=None, lindblad=True, carve_pulse=False, pulse_args=None, filter_width=0.14,
                          pulse_file=None):
        """Full time dynamics for one pulse (reference ``:80-118``); returns ``t.real`` and the outputs."""
        p = ChirpedPulse(tau_0=tau, e_start=detuning, alpha=0, e0=area, polar_x=1.0, t0=4 * tau)
        if tend is None:
            tend = np.round(10 / self.gamma_e) + 100
        self._ensure_pt()
        kw = dict(self.options)
        if carve_pulse and pulse_file is None:
            # a Gaussian carved out of a broader spectrum by a band pass with soft edges (reference :94-98)
            shaped = self._carved_pulse(area, 100, np.round(10 / self.gamma_e), pulse_args or {"width_t": 4, "central_f": 0},
                                        filter_width, 0.01)
            pulse_file, _ = shaped.generate_pulsefiles(suffix="timedynamics", temp_dir=self.options.get("temp_dir", ""),
                                                       in_memory=True)
        if pulse_file is not None:
            kw[self._pulse_file_kw()] = pulse_file
        res = self._system(0, tend, p, lindblad=lindblad, **kw)
        if save:
            export_csv(path + "timedynamics_{:.2f}ps_{:.2f}pi.csv".format(tau, area), res[0].real,
                       *[r.real for r in res[2:]])
        return (res[0].real,) + tuple(res[1:])

    @staticmethod
    def _carved_pulse(area, t0, t_window, pulse_args, filter_width, rise_f):
        """``PulseGenerator`` holding one spectrally carved pulse (reference ``rabi_rotations.py:94-97,176-183``)."""
        from pyaceqd_b200.pulsegenerator import PulseGenerator
        pulse = PulseGenerator(0, t_window, 0.02)
        pulse.add_gaussian_time(t0=t0, sig_or_fwhm='fwhm', field_or_intesity='int', area_time=area, **pulse_args)
        pulse.add_filter_double_erf(central_f=0, width_f=filter_width, rise_f=rise_f)
        pulse.apply_frequency_filter()
        return pulse

    def _filename(self, path, carve_pulse, pulse_args, filter_width):
        name = path + self.prefix
        if carve_pulse:
            name += "carve_{:.2f}ps_{:.3f}nm_".format(pulse_args["width_t"], filter_width)
        if self.phonons:
            name += "{:.1f}K_tau_{:.1f}ps_ae_{:.1f}_ah_{:.2f}_coupl_{:.1f}".format(
                self.temperature, self.tau, self.ae, self.ah_ratio, self.phonon_factor)
        return name

    def get_rabi_rotations(self, detuning=0, integrate=True, plot=False, delete_pt=True, path="", workers=15,
                           carve_pulse=False, pulse_args={"width_t": 4, "central_f": 0}, filter_width=0.14,
                           rise_f=0.01, exp_data=None, plot_dynamic=False, pulse_files=None):
        """``integrate``: occupation integrated over the decay (``gamma_e * int x dt``, photon counts);
        otherwise the occupation right after the pulse (window ``8 tau``, no decay).  Results are cached
        as ``<path><prefix>....csv`` and reused when present (reference ``:120-228``)."""
        filename = self._filename(path, carve_pulse, pulse_args, filter_width)
        if os.path.exists(filename + ".csv"):
            data = np.loadtxt(filename + ".csv", delimiter=",")
            return (data[:, 0], data[:, 1]) if self.n_columns == 1 else (data[:, 0],) + tuple(data[:, 1:].T)
        t_end_add = 0
        if carve_pulse and pulse_files is None:
            # one carved pulse per area; the area that survives the filter replaces the nominal one (reference :176-187)
            t_end_add = 400
            pulse_files = []
            self.areas = np.array(self.areas, dtype=float)
            for i in range(len(self.areas)):
                shaped = self._carved_pulse(self.areas[i], 200, np.round(10 / self.gamma_e), pulse_args, filter_width, rise_f)
                pulse_files.append(shaped.generate_pulsefiles(suffix=str(i), temp_dir=self.options.get("temp_dir", ""),
                                                              in_memory=True)[0])
                self.areas[i] = np.sqrt(shaped.pulse_power)
        self._ensure_pt()
        with BatchExecutor(max_workers=workers) as ex:
            futs = []
            for i, area in enumerate(self.areas):
                p = ChirpedPulse(tau_0=self.tau, e_start=detuning, alpha=0, e0=area, polar_x=1.0, t0=4 * self.tau)
                kw = dict(self.options)
                if pulse_files is not None:
                    kw[self._pulse_file_kw()] = pulse_files[i]
                if integrate:
                    tend = np.round(self.integrate_window_factor / self.gamma_e) + self.integrate_extra + t_end_add
                    futs.append(ex.submit(self._system, 0, tend, p, lindblad=True, suffix=i, **kw))
                else:
                    futs.append(ex.submit(self._system, 0, 8 * self.tau + t_end_add, p, lindblad=False, suffix=i, **kw))
            wait(futs)
        results = np.zeros((self.n_columns, len(self.areas)))
        for i, f in enumerate(futs):
            results[:, i] = self._read(f.result(), integrate)
        export_csv(filename + ".csv", self.areas, *results)
        if delete_pt:
            self.delete_pt_files()
        return (self.areas, results[0]) if self.n_columns == 1 else (self.areas, results)


class RabiRotations(_AreaSweep):
    def __init__(self, dt=0.1, tau=5, area_max=30, n_area=150, gamma_e=1 / 100, phonons=False, temperature=4, ae=5,
                 ah_ratio=1.15, J_from_file=None, phonon_factor=1, t_mem=10, temp_dir=temp_dir) -> None:
        super().__init__(dt, tau, area_max, n_area, gamma_e, phonons, temperature, ae, ah_ratio, J_from_file,
                         phonon_factor, t_mem, temp_dir)

    def _system(self, *a, **kw):
        return tls(*a, **kw)

    def _generate_kwargs(self):
        return {"factor_ah": self.ah_ratio, "phonon_factor": self.phonon_factor, "J_file": self.J_from_file}

    def _read(self, res, integrate):
        t, g, x, pgx, pxg = res
        return self.gamma_e * np.trapezoid(np.real(x), np.real(t)) if integrate else np.real(x[-1])

    def get_J_omega(self, plot=False, n=2000, w_max=15.0):
        """Phonon spectral density of the configured dot (reference ``:37-65`` dumps it through ACE's
        ``Boson_J_print F 0 15 2000``; here it is evaluated directly)."""
        from pyaceqd_b200.pt_builder import qd_phonon_spectral_density
        omega = np.linspace(0.0, w_max, n)
        return omega, qd_phonon_spectral_density(omega, a_e=self.ae, a_h=self.ae / self.ah_ratio)
