"""Pulse shaper: fields built and filtered in the time / frequency domain, fed to the engine as sampled drive tables.

Field path of the reference's ``PulseGenerator`` (``pyaceqd/pulsegenerator.py``: pulse builders ``:108-162``, the
time/frequency bookkeeping ``:284-314``, filters ``:316-433``, ``apply_frequency_filter`` / ``apply_temporal_filter``
``:518-539``, ``generate_pulsefiles`` / ``get_temporal_representation`` ``:1126-1142``).  Same constructor, method
names, argument meaning and units (time in ps, frequencies in THz relative to the rotating frame, ``unit`` = 'Hz' |
'meV' | 'nm'); plotting, the SLM / spectrometer models and the pickle helpers are out of scope.

Both polarisation components live in one ``[2, n]`` array per domain.  ``generate_pulsefiles`` writes the reference's
files (``t Re Im``, ``%.8f``); with ``in_memory=True`` nothing touches the disk: the returned names are keys of an
in-process registry that ``system_ace_stream(pulse_file_x=..., pulse_file_y=...)`` resolves to drive tables, so a
shaped pulse goes from ``get_temporal_representation()`` straight to the operator builder's table interpolation.
"""
from __future__ import annotations

import copy
import math
from typing import Dict, Tuple

import numpy as np
from scipy.special import erf

from pyaceqd_b200 import pulses
from pyaceqd_b200.tools import export_csv

hbar = 0.6582173               # meV ps -- the pulse shaper's own constant (reference pulsegenerator.py:15), NOT constants.hbar
C_NM_THZ = 299792.458          # speed of light in nm THz

# pulse "files" that never touch the disk: name -> (t, values)
_MEMORY_FILES: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}


def memory_file(name: str):
    """The sampled field registered under ``name`` by ``generate_pulsefiles(in_memory=True)`` (or None)."""
    return _MEMORY_FILES.get(name)


def _which(pol: str) -> Tuple[bool, bool]:
    """Components a polarisation selector addresses: 'x', 'y' or 'b'(oth)."""
    c = pol.lower()[0]
    return c in ("b", "x"), c in ("b", "y")


class PulseGenerator:
    def __init__(self, t0, tend=100, dt=0.5, central_wavelength=800, calibration_file=None, f0=None, fend=None, fN=1024,
                 unit='nm') -> None:
        self.calibration_file = calibration_file
        if calibration_file is None:
            self.central_wavelength = central_wavelength
        else:                                          # the rotating frame sits on the measured exciton line
            self._read_calibration_file(calibration_file)
        self.t0 = t0
        if f0 is not None and fend is not None:        # grid chosen from the spectral window
            self.dt = np.abs(1 / (self._Units(fend, unit) - self._Units(f0, unit)))
            self.tend = fN * self.dt + self.t0
        else:
            self.tend, self.dt = tend, dt
        self.time = np.arange(self.t0, self.tend + self.dt, self.dt)
        n = len(self.time)
        self.frequencies = -np.fft.fftshift(np.fft.fftfreq(n, d=self.dt))      # negative: rotating frame
        self.df = np.abs(self.frequencies[0] - self.frequencies[1])
        self.angular_frequencies = 2 * np.pi * self.frequencies
        self.energies = 2 * np.pi * hbar * self.frequencies
        self.central_frequency = C_NM_THZ / self.central_wavelength
        self.central_energy = self.central_frequency * hbar * 2 * np.pi
        self.wavelengths = C_NM_THZ / (self.central_frequency + self.frequencies)
        self._field_t = np.zeros((2, n), dtype=complex)       # [x|y, sample]
        self._field_f = np.zeros((2, n), dtype=complex)
        self._filter_f = np.zeros((2, n), dtype=complex)
        self._filter_t = np.ones((2, n), dtype=complex)
        self.pulse_power = 0
        self.action_counter = 0

    # the reference's attribute names, as views of the packed arrays
    temporal_representation_x = property(lambda s: s._field_t[0])
    temporal_representation_y = property(lambda s: s._field_t[1])
    frequency_representation_x = property(lambda s: s._field_f[0])
    frequency_representation_y = property(lambda s: s._field_f[1])
    frequency_filter_x = property(lambda s: s._filter_f[0])
    frequency_filter_y = property(lambda s: s._filter_f[1])
    temporal_filter_x = property(lambda s: s._filter_t[0])
    temporal_filter_y = property(lambda s: s._filter_t[1])

    # ------------------------------------------------------------------ units and small helpers
    def _read_calibration_file(self, calibration_file):
        """Emission lines of a measured quantum dot (INI file, the one ``tools.read_calibration_file`` reads) as
        THz offsets from the exciton line, which becomes the rotating frame (reference ``:66-86``)."""
        import configparser
        config = configparser.ConfigParser()
        config.read(calibration_file)
        self.central_wavelength = float(config['EMISSION']['exciton_wavelength'])
        self.biexciton_wavelength = float(config['EMISSION']['biexciton_wavelength'])
        self.dark_wavelength = float(config['EMISSION']['dark_wavelength'])
        self.fss_bright = float(config['SPLITTING']['fss_bright'])
        self.fss_dark = float(config['SPLITTING']['fss_dark'])
        self.lifetime_exciton = float(config['LIFETIMES']['exciton'])
        self.lifetime_biexciton = float(config['LIFETIMES']['biexciton'])
        x, b, d = (self._Units(w, 'nm') for w in (self.central_wavelength, self.biexciton_wavelength, self.dark_wavelength))
        half_bright = self._Units(self.fss_bright * 1e-3 / 2, 'mev')
        half_dark = self._Units(self.fss_dark * 1e-3 / 2, 'mev')
        self.exciton_x_emission, self.exciton_y_emission = x + half_bright, x - half_bright
        self.biexciton_x_emission, self.biexciton_y_emission = b - half_bright, b + half_bright
        self.dark_x_emission, self.dark_y_emission = d + half_dark, d - half_dark
        self.tpe_resonance = (x + b) / 2

    def _Units(self, value, unit='Hz'):
        """meV or nm -> THz offset from the rotating frame ('nm' accepts absolute or relative wavelengths)."""
        u = unit.lower()[0]
        if u == 'm':
            return value / (2 * np.pi * hbar)
        if u == 'n':
            if np.abs(value - self.central_wavelength) < np.abs(value):
                value = value - self.central_wavelength
            return C_NM_THZ / (self.central_wavelength + value) - C_NM_THZ / self.central_wavelength
        return value

    def _Units_inverse(self, value, unit='Hz'):
        u = unit.lower()[0]
        if u == 'm':
            return value * (2 * np.pi * hbar)
        if u == 'n':
            return C_NM_THZ / (C_NM_THZ / self.central_wavelength + value)
        return value

    @staticmethod
    def _Taylor(frequency, frequency_0=0, coefficients=()):
        phase = np.zeros_like(frequency)
        for k, c in enumerate(coefficients):
            phase += c / math.factorial(k) * (frequency - frequency_0) ** k
        return phase

    @staticmethod
    def _sig_fwhm(field_int, sig_fwhm, width):
        """Width given as sigma or FWHM of the field or of the intensity -> sigma of the field."""
        field = field_int.lower()[0] == 'f'
        sigma = sig_fwhm.lower()[0] == 's'
        if field:
            return width if sigma else width / (2 * np.sqrt(np.log(2) * 2))
        return np.sqrt(2) * width if sigma else width / (2 * np.sqrt(np.log(2)))

    @staticmethod
    def _sigmoid(x, center, width, rise):
        return 1 / (1 + np.exp(-(x - (center - width / 2)) / rise)) / (1 + np.exp(-((center + width / 2) - x) / rise))

    @staticmethod
    def _normalise_polarisation(pol):
        pol = np.array(pol, dtype=complex)
        return tuple(pol / np.sqrt(np.abs(pol[0] ** 2) + np.abs(pol[1] ** 2)))

    def _set_action_counter(self, action_counter):
        self.action_counter = action_counter

    def _add_action_counter(self, n=1):
        self.action_counter += n

    def _update_pulse_power(self):
        self.pulse_power = np.trapezoid(y=np.sum(np.abs(self._field_t) ** 2, axis=0), x=np.real(self.time))

    # ------------------------------------------------------------------ building pulses
    def _add_time(self, pulse_x_time, pulse_y_time):
        add = np.array([pulse_x_time, pulse_y_time], dtype=complex)
        self._field_t += add
        self._field_f += np.fft.fftshift(np.fft.fft(add, axis=1), axes=1)
        self._update_pulse_power()
        self._add_action_counter()

    def _add_spectral(self, pulse_x_freq, pulse_y_freq):
        add = np.array([pulse_x_freq, pulse_y_freq], dtype=complex)
        self._field_f += add
        self._field_t += np.fft.ifft(np.fft.ifftshift(add, axes=1), axis=1)
        self._update_pulse_power()
        self._add_action_counter()

    def add_gaussian_time(self, width_t, central_f=0, alpha=0, t0=0, area_time=1, polarisation=[1, 0], phase=0,
                          field_or_intesity='field', sig_or_fwhm='sig', unit='Hz'):
        """Gaussian in time: ``area_time`` = transform-limited pulse area, ``alpha`` = chirp in ps^2."""
        energy = self._Units(central_f, unit) * hbar * 2 * np.pi
        tau = np.abs(self._sig_fwhm(field_or_intesity, sig_or_fwhm, width_t))
        px, py = self._normalise_polarisation(polarisation)
        field = pulses.ChirpedPulse(tau, energy, alpha, t0, area_time, px, phase).get_total(self.time)
        self._add_time(field * px, field * py)

    def add_gaussian_freq(self, width_f, central_f=0, area_time=1, polarisation=[1, 0], field_or_intesity='field',
                          sig_or_fwhm='sig', phase_taylor=[], shift_time=0, unit='Hz'):
        """Gaussian in frequency with a spectral phase given by its Taylor coefficients (ps^n)."""
        f_c = self._Units(central_f, unit)
        sig = self._sig_fwhm(field_or_intesity, sig_or_fwhm, np.abs(self._Units(width_f, unit)))
        px, py = self._normalise_polarisation(polarisation)
        spec = 1 / self.dt * area_time * np.exp(-(self.frequencies - f_c) ** 2 / (2 * sig ** 2)) * np.exp(
            1j * self._Taylor(self.frequencies * 2 * np.pi, f_c * 2 * np.pi, coefficients=phase_taylor))
        spec = spec * np.exp(1j * 2 * np.pi * self.frequencies * (shift_time - np.min(self.time)))
        self._add_spectral(spec * px, spec * py)

    def add_rectangle_frequ(self, central_f, width_f, hight, phase_taylor=[], polarisation=[1, 0], shift_time=0, unit='Hz'):
        f_c, w = self._Units(central_f, unit), np.abs(self._Units(width_f, unit))
        px, py = self._normalise_polarisation(polarisation)
        spec = np.zeros_like(self.frequencies, dtype=complex)
        spec[np.abs(self.frequencies - f_c) <= w / 2] = hight
        spec = spec * np.exp(1j * self._Taylor(self.frequencies * 2 * np.pi, f_c * 2 * np.pi, coefficients=phase_taylor))
        spec = spec * np.exp(1j * 2 * np.pi * self.frequencies * (shift_time - np.min(self.time)))
        self._add_spectral(spec * px, spec * py)

    # ------------------------------------------------------------------ filters
    def _add_filter(self, filter, pol='both', merging='+', cap_transmission=True):
        """Merge a transmission function into the frequency filter: '+' add, '*' multiply, 'm' overlay (maximum)."""
        for k, on in enumerate(_which(pol)):
            if not on:
                continue
            if merging == '+':
                self._filter_f[k] += filter
            elif merging == '*':
                self._filter_f[k] *= filter
            elif merging.lower()[0] == 'm':
                # element-wise np.max([value, new]) of the reference: complex numbers order by real, then imaginary part
                cur = self._filter_f[k]
                new = np.asarray(filter, dtype=complex)
                take = (new.real > cur.real) | ((new.real == cur.real) & (new.imag > cur.imag))
                self._filter_f[k] = np.where(take, new, cur)
        if cap_transmission and np.any(np.abs(self._filter_f) > 1):
            self._filter_f[self._filter_f > 1] = 1

    def _add_filter_time(self, filter, pol='both', merging='+', cap_transmission=True):
        for k, on in enumerate(_which(pol)):
            if not on:
                continue
            if merging == '+':
                self._filter_t[k] += filter
            elif merging == '*':
                self._filter_t[k] *= filter
        if cap_transmission and np.any(np.abs(self._filter_t) > 1):
            self._filter_t[self._filter_t > 1] = 1

    def add_filter_rectangle(self, central_f=None, width_f=None, transmission=1, cap_transmission=True, polarisation='b',
                             invert=False, merging='+', unit='Hz'):
        if central_f is None:
            filt = np.ones_like(self.frequencies, dtype=complex) * transmission
        else:
            f_c, w = self._Units(central_f, unit), np.abs(self._Units(width_f, unit))
            filt = np.zeros_like(self.frequencies, dtype=complex)
            filt[np.abs(self.frequencies - f_c) <= w / 2] = transmission
            if invert:
                filt = 1 - filt
        self._add_filter(filt, polarisation, merging=merging, cap_transmission=cap_transmission)

    def add_filter_gaussian(self, central_f, width_f, transmission=1, super_gauss=1, polarisation='b', field_int='field',
                            sig_fwhm='sig', invert=False, merging='+', unit='Hz', phase=False):
        f_c = self._Units(central_f, unit)
        sig = self._sig_fwhm(field_int, sig_fwhm, np.abs(self._Units(width_f, unit)))
        g = np.exp(-((self.frequencies - f_c) ** 2 / (2 * sig ** 2)) ** super_gauss) * transmission
        if invert:
            g = 1 - g
        if phase:
            self._add_filter(np.exp(1j * g * np.pi * 2. * transmission), polarisation, merging='*')
        else:
            self._add_filter(g, polarisation, merging=merging)

    def add_filter_sigmoid(self, central_f, width_f, rise_f, transmission=1, polarisation='b', invert=False, merging='+',
                           unit='Hz'):
        s = self._sigmoid(self.frequencies, self._Units(central_f, unit), np.abs(self._Units(width_f, unit)),
                          np.abs(self._Units(rise_f, unit)))
        s = s / np.max(s) * transmission
        self._add_filter(1 - s if invert else s, polarisation, merging)

    def add_filter_double_erf(self, central_f, width_f, rise_f, transmission=None, polarisation='b', invert=False,
                              merging='+', unit='Hz', cap_transmission=True, field_int='int', sig_fwhm='fwhm'):
        """Band pass with Gaussian-broadened edges (step convolved with a Gaussian of width ``rise_f``)."""
        f_c, w = self._Units(central_f, unit), np.abs(self._Units(width_f, unit))
        rise = self._sig_fwhm(field_int, sig_fwhm, np.abs(self._Units(rise_f, unit))) * np.sqrt(2)
        filt = 0.5 * (erf((self.frequencies - f_c + w / 2) / rise) - erf((self.frequencies - f_c - w / 2) / rise))
        if transmission is not None:
            filt = filt / np.max(filt) * transmission
        self._add_filter(1 - filt if invert else filt, polarisation, merging, cap_transmission=cap_transmission)

    def add_phase_filter(self, central_f=0, phase_taylor=[], polarisation='b', unit='Hz', f_start=None, f_end=None):
        lo = np.min(self.frequencies) if f_start is None else self._Units(f_start, unit)
        hi = np.max(self.frequencies) if f_end is None else self._Units(f_end, unit)
        ph = self._Taylor(self.frequencies * 2 * np.pi, self._Units(central_f, unit) * 2 * np.pi, coefficients=phase_taylor)
        ph[(self.frequencies < lo) | (self.frequencies > hi)] = 0
        self._add_filter(np.exp(1j * ph), pol=polarisation, merging='*')

    def apply_frequency_filter(self, pol='b'):
        """Multiply the spectrum by the frequency filter and refresh the time-domain field."""
        c = pol.lower()[0]
        # (operator precedence of the reference: 'b' always applies; 'x' / 'y' only if that component is non-zero)
        for k, name in enumerate("xy"):
            if c == 'b' or (c == name and np.any(self._field_f[k] != 0)):
                self._field_f[k] *= self._filter_f[k]
                self._field_t[k] = np.fft.ifft(np.fft.ifftshift(self._field_f[k]))
        self._update_pulse_power()
        self._add_action_counter()

    def apply_temporal_filter(self, pol='b'):
        c = pol.lower()[0]
        for k, name in enumerate("xy"):
            if c == 'b' or (c == name and np.any(self._field_t[k] != 0)):
                self._field_t[k] *= self._filter_t[k]
                self._field_f[k] = np.fft.fftshift(np.fft.fft(self._field_t[k]))
        self._update_pulse_power()
        self._add_action_counter()

    def set_pulse_power(self, power):
        if self.pulse_power == 0:
            print('Initial pulse power is 0.')
            return
        self.clear_filter()
        self.add_filter_rectangle(transmission=np.sqrt(power / self.pulse_power), cap_transmission=False)
        self.apply_frequency_filter()
        self.clear_filter()

    def set_rotating_frame(self, new_rf=None, unit='nm'):
        if isinstance(new_rf, str):
            self._read_calibration_file(new_rf)
        else:
            self.central_wavelength = self._Units_inverse(self._Units(new_rf, unit), 'nm')
        new_f = C_NM_THZ / self.central_wavelength
        self.central_energy = new_f * hbar * 2 * np.pi
        self._field_t *= np.exp(-1j * 2 * np.pi * (self.central_frequency - new_f) * self.time)
        self._field_f = np.fft.fftshift(np.fft.fft(self._field_t, axis=1), axes=1)
        self.central_frequency = new_f
        self.wavelengths = C_NM_THZ / (self.central_frequency + self.frequencies)

    # ------------------------------------------------------------------ hand-over to the propagation engine
    def get_temporal_representation(self, abs_only=False):
        if abs_only:
            return self.time, np.abs(self._field_t[0]), np.abs(self._field_t[1])
        return self.time, self._field_t[0], self._field_t[1]

    def generate_pulsefiles(self, temp_dir='', file_name='pulse_time', suffix='', abs_only=False, precision=8,
                            in_memory=False):
        """Pulse files ``t Re Im`` of both polarisations for ``system(..., pulse_file_x=, pulse_file_y=)``
        (reference ``:1126-1137``).  ``in_memory=True``: no files; the names resolve inside this process."""
        names = [temp_dir + file_name + str(suffix) + '_x.dat', temp_dir + file_name + str(suffix) + '_y.dat']
        for name, field in zip(names, self._field_t):
            re, im = (np.abs(field), np.zeros(len(field))) if abs_only else (np.real(field), np.imag(field))
            if in_memory:
                q = 10.0 ** precision        # the same quantisation a %.8f file would apply
                _MEMORY_FILES[name] = (np.array(self.time, dtype=float), np.round(re * q) / q + 1j * np.round(im * q) / q)
            else:
                export_csv(name, self.time, re, im, precision=precision, delimit=' ')
        return names[0], names[1]

    def merge_pulses(self, input_pulse):
        """Add another generator's field, interpolated onto this time grid (cubic, zero outside)."""
        from scipy import interpolate
        other = input_pulse.copy_pulse()
        if other.central_wavelength != self.central_wavelength:
            print('Caution MERGING: Central wavelength of pulses do not agree!')
            other.set_rotating_frame(self.central_wavelength)
        if other.dt != self.dt:
            print('CAUTION MERGING: Time steps of pulses do not agree!')
        comp = []
        for k in range(2):
            parts = [interpolate.interp1d(other.time, part(other._field_t[k]), kind='cubic', fill_value=0, bounds_error=False)
                     for part in (np.real, np.imag)]
            comp.append(parts[0](self.time) + 1j * parts[1](self.time))
        self._add_time(comp[0], comp[1])

    # ------------------------------------------------------------------ housekeeping
    def clear_all(self):
        self.clear_filter()
        self.clear_pulses()
        self._set_action_counter(0)

    def clear_filter(self):
        self._filter_f = np.zeros_like(self._filter_f)
        self._filter_t = np.ones_like(self._filter_t)

    def clear_pulses(self):
        self._field_t = np.zeros_like(self._field_t)
        self._field_f = np.zeros_like(self._field_f)

    def copy_pulse(self):
        return copy.deepcopy(self)
