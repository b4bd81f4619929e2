"""Analytic drive fields.  Same public classes, constructor arguments and method names as the
reference's ``pyaceqd/pulses.py`` (``Pulse``, ``AsymmetricPulse``, ``ChirpedPulse``,
``PulseTrain``, ``CWLaser``, ``SmoothRectangle``), so user scripts keep working; the
implementation is organised around one shared Gaussian helper.

Field convention (reference ``pulses.py:82-83``): ``get_total(t) = envelope(t) * exp(-i*phi(t))``
with ``phi(t) = (e_start/hbar)(t-t0) + w_gain/2 (t-t0)^2 + phase`` (``:64-68``) and the envelope
normalised such that its time integral is ``e0`` (``:38-39``) -- a pulse area of ``pi*e0`` after
the ``-0.5*pi*hbar`` prefactor of ``general_system.py:279``.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf

from . import constants

hbar = constants.hbar  # meV*ps


def _gauss(t, t0, sigma):
    z = (np.asarray(t, dtype=float) - t0) / sigma
    return np.exp(-0.5 * z * z)


class Pulse:
    """Gaussian pulse with linear chirp rate ``w_gain`` (1/ps^2) and carrier ``e_start`` (meV)."""

    def __init__(self, tau, e_start, w_gain=0, t0=0, e0=1, phase=0, polar_x=1, polars=None):
        self.tau = tau
        self.e_start = e_start
        self.w_gain = float(w_gain)
        self.t0 = t0
        self.e0 = e0
        self.phase = phase
        self.freq = None     # optional callable overriding the instantaneous frequency
        self.phase_ = None   # optional callable overriding the full phase
        if polars is not None:
            nrm = np.sqrt(abs(polars[0]) ** 2 + abs(polars[1]) ** 2)
            self.polar_x, self.polar_y = polars[0] / nrm, polars[1] / nrm
        else:
            self.polar_x, self.polar_y = polar_x, np.sqrt(1 - polar_x ** 2)

    def __repr__(self):
        return "{}(tau={!r}, e_start={!r}, w_gain={!r}, t0={!r}, e0={!r})".format(
            type(self).__name__, self.tau, self.e_start, self.w_gain, self.t0, self.e0)

    # -- carrier ---------------------------------------------------------------------------
    def get_energy(self):
        return self.e_start, self.w_gain

    def set_energy(self, e_start, w_gain):
        self.e_start, self.w_gain = e_start, w_gain

    def set_frequency(self, f):
        self.freq = f

    def set_phase(self, f):
        self.phase_ = f

    def get_frequency(self, t):
        """Instantaneous angular frequency d(phi)/dt."""
        if self.freq is not None:
            return self.freq(t)
        return self.e_start / hbar + self.w_gain * (t - self.t0)

    def get_full_phase(self, t):
        if self.phase_ is not None:
            return self.phase_(t)
        s = t - self.t0
        return (self.e_start / hbar) * s + 0.5 * self.w_gain * s * s + self.phase

    def get_energies(self):
        """Energy sweep (meV) of the chirp between -tau and +tau."""
        return abs(self.get_frequency(self.tau) - self.get_frequency(-self.tau)) * hbar

    # -- envelope --------------------------------------------------------------------------
    def get_envelope(self, t):
        return self.e0 * _gauss(t, self.t0, self.tau) / (np.sqrt(2 * np.pi) * self.tau)

    def get_integral(self, t):
        return 0.5 * self.e0 * (1 - erf((self.t0 - t) / (np.sqrt(2) * self.tau)))

    def get_total(self, t):
        return self.get_envelope(t) * np.exp(-1j * self.get_full_phase(t))

    def copy(self):
        return Pulse(self.tau, self.e_start, self.w_gain, self.t0, self.e0, self.phase, self.polar_x)


class AsymmetricPulse(Pulse):
    """Gaussian with width ``tau1`` before ``t0`` and ``tau2`` after, continuous at ``t0``."""

    def __init__(self, tau1, tau2, e_start, t0=0, e0=1, phase=0, polar_x=1, polars=None):
        self.tau1, self.tau2 = tau1, tau2
        super().__init__(tau1, e_start, w_gain=0, t0=t0, e0=e0, phase=phase, polar_x=polar_x, polars=polars)

    def get_envelope(self, t):
        t = np.asarray(t, dtype=float)
        sigma = np.where(t <= self.t0, self.tau1, self.tau2)
        # both halves share the tau1 normalisation so the envelope is continuous
        return self.e0 * _gauss(t, self.t0, sigma) / (np.sqrt(2 * np.pi) * self.tau1)

    def copy(self):
        return AsymmetricPulse(self.tau1, self.tau2, self.e_start, self.t0, self.e0, self.phase, self.polar_x)


class ChirpedPulse(Pulse):
    """Transform-limited Gaussian of width ``tau_0`` stretched by a chirp ``alpha`` (ps^2)."""

    def __init__(self, tau_0, e_start, alpha=0, t0=0, e0=1 * np.pi, polar_x=1, phase=0, polars=None):
        self.tau_0, self.alpha = tau_0, alpha
        super().__init__(tau=np.sqrt(alpha ** 2 / tau_0 ** 2 + tau_0 ** 2), e_start=e_start,
                         w_gain=alpha / (alpha ** 2 + tau_0 ** 4), t0=t0, e0=e0, polar_x=polar_x,
                         phase=phase, polars=polars)

    def get_parameters(self):
        return "tau: {:.4f} ps , a: {:.4f} ps^-2".format(self.tau, self.w_gain)

    def get_envelope(self, t):
        return self.e0 * _gauss(t, self.t0, self.tau) / np.sqrt(2 * np.pi * self.tau * self.tau_0)

    def get_integral(self, t):
        return super().get_integral(t) * self.get_ratio()

    def get_ratio(self):
        """Pulse-area ratio chirped / unchirped."""
        return np.sqrt(self.tau / self.tau_0)

    def copy(self):
        return ChirpedPulse(self.tau_0, self.e_start, self.alpha, self.t0, self.e0, self.polar_x, self.phase)


class PulseTrain:
    """``n_pulses`` repetitions (spacing ``delta_t``) of a group of pulses."""

    def __init__(self, delta_t, n_pulses, *pulses, t_shift=0):
        self.delta_t, self.n_pulses = delta_t, n_pulses
        self.pulses = list(pulses)
        self.t_shift = t_shift

    def _shifted(self, t):
        for i in range(self.n_pulses):
            yield t - self.delta_t * i - self.t_shift

    def get_total(self, t):
        acc = np.zeros_like(t, dtype=complex)
        for ts in self._shifted(t):
            for p in self.pulses:
                acc = acc + p.get_total(ts)
        return acc

    def get_total_xy(self, t):
        fx = np.zeros_like(t, dtype=complex)
        fy = np.zeros_like(t, dtype=complex)
        for ts in self._shifted(t):
            for p in self.pulses:
                f = p.get_total(ts)
                fx = fx + p.polar_x * f
                fy = fy + p.polar_y * f
        return fx, fy


class CWLaser(Pulse):
    """Continuous-wave drive of constant amplitude ``e0`` (no switch-on)."""

    def __init__(self, e0, e_start=0, polar_x=1, phase=0, polars=None):
        super().__init__(tau=5, e_start=e_start, e0=e0, polar_x=polar_x, polars=polars, phase=phase)

    def get_envelope(self, t):
        return self.e0

    def copy(self):
        return CWLaser(self.e0, self.e_start, self.polar_x, self.phase)


class SmoothRectangle(Pulse):
    """Rectangle of length ``tau`` centred at ``t0`` with sigmoid edges of rate ``1/alpha_onoff``."""

    def __init__(self, tau, e_start, w_gain=0, t0=0, e0=1, phase=0, alpha_onoff=0.1, polar_x=1, polars=None):
        self.alpha_onoff = alpha_onoff
        self.alpha = 1 / alpha_onoff
        super().__init__(tau, e_start, w_gain=w_gain, t0=t0, e0=e0, phase=phase, polar_x=polar_x, polars=polars)

    def get_envelope_f(self):
        return lambda t: self.get_envelope(t)

    def get_envelope(self, t):
        rise = 1 + np.exp(-self.alpha * (t + self.tau / 2 - self.t0))
        fall = 1 + np.exp(-self.alpha * (-t + self.tau / 2 + self.t0))
        return self.e0 / (rise * fall)

    def copy(self):
        return SmoothRectangle(self.tau, self.e_start, self.w_gain, self.t0, self.e0, self.phase,
                               self.alpha_onoff, self.polar_x)
