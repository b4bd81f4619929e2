"""``timebin.onephoton.OnePhotonCavity`` and the ``two_level_system.reduced_dark`` sweeps.

The diagonal bookkeeping of ``g1_t1`` is checked with a stand-in system whose outputs encode (row time, operator
time); the physics runs on the oracle backend (a 3 x 3 emitter-cavity system has NL = 81)."""
import numpy as np

from oracle_backend import oracle_backend
from pyaceqd_b200.pulses import ChirpedPulse


class Recorder:
    """system(t0, tend, ..., multitime_op=, output_ops=) -> [t, f(t, t_mto), g(t, t_mto)] on the dt grid."""

    def __init__(self, dt):
        self.dt, self.calls = dt, []

    @staticmethod
    def value(t, t_mto):
        return np.sin(0.7 * t) + 1j * np.cos(0.3 * t_mto) + 0.01 * t * t_mto

    def __call__(self, t0, tend, *pulses, multitime_op=None, output_ops=None, dt=None, **kw):
        n = int(round((tend - t0) / self.dt))
        t = t0 + self.dt * np.arange(n + 1)
        t_mto = multitime_op["time"]
        self.calls.append((tend, t_mto))
        return np.array([t, 2.0 * self.value(t, t_mto), self.value(t, t_mto)])


def _cavity(system, tmp, dt=0.25, tb=1.0):
    from pyaceqd_b200.timebin.onephoton import OnePhotonCavity
    p = ChirpedPulse(tau_0=0.3, e_start=0, alpha=0, t0=1.0, e0=3.0, polar_x=0)
    return OnePhotonCavity(system, p, dt=dt, tb=tb, t_simul=12, options={"temp_dir": tmp})


def test_g1_t1_fills_every_cell_of_the_t1_t2_grid(tmp_path):
    """Cell (a, b) must hold the output at ``t1[a]`` of a run whose operator acts at ``t1[a] + t2[b] - T_sep``
    (reference onephoton.py:189-271): rising, full and falling anti-diagonals together cover the grid once.  The
    reference's run lengths and operator times are consistent for a square grid (``tend - t0 == 2 tb``) only; other
    shapes are filled exactly as it fills them."""
    rec = Recorder(0.25)
    cav = _cavity(rec, str(tmp_path) + "/")
    t1, g1 = cav.g1_t1(t0=3, tend=5, T_sep=1)
    t2 = np.linspace(-1.0, 1.0, 9)
    want = np.trapezoid(Recorder.value(t1[:, None], np.round(t1[:, None] + t2[None, :] - 1, 3)), t2, axis=1)
    assert np.allclose(t1, np.linspace(3, 5, 9)) and np.abs(g1 - want).max() < 1e-12
    assert len(rec.calls) == len(t1) + len(t2) - 1            # one run per anti-diagonal


def test_g1_t1t2_and_g1_t1t_windows(tmp_path):
    rec = Recorder(0.25)
    cav = _cavity(rec, str(tmp_path) + "/")
    t1, g = cav.g1_t1t2(t0=2, tend=3, T_sep=0.5)
    t2 = np.linspace(-1.0, 1.0, 9)
    for i, t in enumerate(t1):
        tm = t - 0.5
        pos = Recorder.value(tm + t2[5:], tm)                  # tau > 0: second output after the operator
        two_sided = np.concatenate([np.conj(pos[::-1]), [2.0 * Recorder.value(tm, tm)], pos])
        assert abs(g[i] - np.trapezoid(two_sided, t2)) < 1e-12
    t1, g = cav.g1_t1t(t0=2, tend=3, T_sep=1.0)
    for i, t in enumerate(t1):
        assert abs(g[i] - np.trapezoid(Recorder.value(t + t2, t - 1.0), t2)) < 1e-12


def test_cavity_emitter_on_the_oracle_backend(tmp_path):
    """Emitter + cavity (reduced_dark.darkmodel_photons, reference :32-53): one batch per method; the cavity photon's
    G1(t, t) integral is real and positive, and the three routes agree where they describe the same quantity."""
    from pyaceqd_b200.two_level_system.reduced_dark import darkmodel_photons
    tmp = str(tmp_path) + "/"
    opts = {"lindblad": True, "temp_dir": tmp, "cav_coupl": 1.0, "cav_loss": 0.5, "delta_cx": 0.0, "rad_loss": 0.1}
    from pyaceqd_b200.timebin.onephoton import OnePhotonCavity
    p = ChirpedPulse(tau_0=0.3, e_start=0, alpha=0, t0=1.0, e0=3.0, polar_x=0)
    with oracle_backend() as eng:
        cav = OnePhotonCavity(darkmodel_photons, p, dt=0.25, tb=1.0, t_simul=12, options=opts)
        t1, a = cav.g1_t1t2(t0=2, tend=4, T_sep=0)
        _, b = cav.g1_t1(t0=3, tend=5, T_sep=1)
        t, g, x, d = darkmodel_photons(0, 5, p, dt=0.25, **opts)
    assert [c[0] for c in eng.calls] == [9, 17, 1]
    # two-sided integral of a conjugate-symmetric function: real; dominated by the photon number around the pulse
    assert np.abs(a.imag).max() < 1e-12 and a.real.max() > 0.5 and np.abs(b).max() > 1e-2
    assert np.abs(g + x + d - 1).max() > 1e-3                 # photons carry population out of the zero-photon block
    # the tau = 0 element is the photon number at the operator time
    mto = {"operator": cav.sigma_xdag, "applyFrom": "_right", "applyBefore": "false", "time": 2.0}
    outs = ["|0><0|_3 otimes |1><1|_3", cav.sigma_x, "Id_3 otimes n_3"]
    with oracle_backend():
        r = darkmodel_photons(0, 3.0, p, dt=0.25, multitime_op=mto, output_ops=outs, **opts)
        plain = darkmodel_photons(0, 3.0, p, dt=0.25, output_ops=outs, **opts)
    assert abs(r[1][8] - plain[1][8]) < 1e-12 and plain[3][8].real > plain[1][8].real > 0


def test_reduced_dark_bin_integrals_and_el_sweeps(tmp_path):
    """reduced_dark.G1_ee / G1_ll / G1_el / G1_easy_el (reference :55-183) on the oracle backend."""
    from pyaceqd_b200.two_level_system.reduced_dark import G1_easy_el, G1_ee, G1_el, G1_ll, darkmodel
    tmp = str(tmp_path) + "/"
    p = ChirpedPulse(tau_0=0.3, e_start=0, alpha=0, t0=1.0, e0=3.0, polar_x=0)
    with oracle_backend() as eng:
        ee, ll = G1_ee(p, dt=0.25, tb=4.0, temp_dir=tmp, delta_xd=0.5), G1_ll(p, dt=0.25, tb=4.0, temp_dir=tmp, delta_xd=0.5)
        t, g, x, d = darkmodel(0, 8.0, p, dt=0.25, delta_xd=0.5, gamma_e=1 / 65, lindblad=True, temp_dir=tmp)
        n0 = len(eng.calls)
        t1, t2, g1 = G1_el(p, dt=0.5, dtau=0.25, tb=4.0, temp_dir=tmp, gaussian_t=2.0, delta_xd=0.5)
        t1b, easy = G1_easy_el(p, dt=0.5, dtau=0.25, tb=4.0, temp_dir=tmp, gaussian_t=2.0, delta_xd=0.5)
        calls = eng.calls[n0:]
    assert abs(ee - np.trapezoid(x.real[:17], t.real[:17])) < 1e-12
    assert abs(ll - np.trapezoid(x.real[-16:], t.real[-16:])) < 1e-12
    assert [c[0] for c in calls] == [len(t1), len(t1)]                       # each sweep is one batch
    assert g1.shape == (len(t1), 17) and np.allclose(t2, np.linspace(0, 4, 17)) and np.allclose(t1, t1b)
    assert easy.shape == (len(t1),) and np.abs(easy).max() > 1e-3
    # every G1_el run ends at 2 tb whatever t1 is: row i equals an eager run with the operator at t1[i]
    with oracle_backend():
        from pyaceqd_b200.pulses import ChirpedPulse as CP
        from pyaceqd_b200.tools import export_csv
        grid = np.arange(0, 2.1 * 4.0, step=0.25)
        f = p.polar_y * p.get_total(grid)
        export_csv(tmp + "y.dat", grid, f.real, f.imag, precision=8, delimit=' ')
        export_csv(tmp + "x.dat", grid, 0 * f.real, 0 * f.real, precision=8, delimit=' ')
        r = darkmodel(0, 8.0, p, dt=0.25, delta_xd=0.5, gamma_e=1 / 65, lindblad=True, temp_dir=tmp,
                      pulse_file_x=tmp + "x.dat", pulse_file_y=tmp + "y.dat",
                      output_ops=["|0><0|_3", "|1><1|_3", "|2><2|_3", "|0><1|_3"],
                      multitime_op={"operator": "|1><0|_3", "applyFrom": "_right", "applyBefore": "false", "time": t1[2]})
    assert abs(g1[2, 0] - r[2][-17]) < 1e-12 and np.abs(g1[2, 1:] - r[4][-16:]).max() < 1e-12
    assert abs(easy[2] - r[4][int(round((t1[2] + 4.0) / 0.25))]) < 1e-12
