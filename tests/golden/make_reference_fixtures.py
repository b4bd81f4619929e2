"""Generate golden fixtures by IMPORTING the reference (run in the build container only;
/root/reference does not exist on the GPU box).  matplotlib / ACEutils are absent here, so they
are stubbed in sys.modules before import.  Output: reference_host.json.

    python tests/golden/make_reference_fixtures.py
"""
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.collections", "ACEutils"):
    sys.modules.setdefault(name, types.ModuleType(name))
for name in ("Parameters", "FreePropagator", "ProcessTensors", "InitialState", "OutputPrinter", "TimeGrid", "Simulation",
             "read_outfile", "DynamicalMap"):          # names general_system.py:14 imports from the absent pybind module
    setattr(sys.modules["ACEutils"], name, None)

import pyaceqd.pulses as rp  # noqa: E402
import pyaceqd.tools as rt  # noqa: E402

out = {}
# --- pulses (reference pulses.py): complex fields on a fixed grid
t = np.linspace(-30.0, 60.0, 181)
specs = {
    "Pulse": ("Pulse", dict(tau=3, e_start=1.2, w_gain=0.05, t0=10, e0=2.0, phase=0.3, polar_x=0.6)),
    "ChirpedPulse": ("ChirpedPulse", dict(tau_0=3, e_start=-2, alpha=20, t0=12, e0=5)),
    "AsymmetricPulse": ("AsymmetricPulse", dict(tau1=2, tau2=6, e_start=0.5, t0=5, e0=1.5)),
    "CWLaser": ("CWLaser", dict(e0=0.3, e_start=1)),
    "SmoothRectangle": ("SmoothRectangle", dict(tau=20, e_start=0.2, t0=15, e0=0.4)),
}
out["pulse_t"] = t.tolist()
out["pulses"] = {}
for key, (cls, kw) in specs.items():
    p = getattr(rp, cls)(**kw)
    f = p.get_total(t)
    out["pulses"][key] = dict(cls=cls, kw=kw, re=np.real(f).tolist(), im=np.imag(f).tolist(),
                              freq=np.asarray(p.get_frequency(t), dtype=float).tolist() if cls != "CWLaser" else None,
                              polar_y=float(p.polar_y))
# --- tools: interval merging cases of tests/test_merge_interval.py:5-23
cases = [[[-1, 1], [1, 2], [5, 8]], [[-1, 2], [1, 3], [4, 8]], [[-1, 3], [1, 3], [4, 8]],
         [[-1, 1], [1, 3], [4, 8]], [[-1, 7], [1, 3], [4, 8]]]
out["merge"] = [dict(inp=c, out=rt._merge_intervals([list(x) for x in c])) for c in cases]
# --- construct_t grids of tests/test_merge_interval.py:25-33 (note the dt_exp positional quirk:
#     the first pulse binds to dt_exp there; we call with dt_exp explicit AND in the quirky form)
p1 = rp.ChirpedPulse(tau_0=1, e_start=0, t0=4)
p2 = rp.ChirpedPulse(tau_0=1, e_start=0, t0=20)
p3 = rp.ChirpedPulse(tau_0=1, e_start=0, t0=5)
out["construct_t"] = {
    "two_pulses": rt.construct_t(0, 80, 0.1, 1.0, None, p1, p2, simple_exp=False).tolist(),
    "quirk_first_pulse_is_dt_exp": rt.construct_t(0, 80, 0.1, 1.0, p1, p2, simple_exp=False).tolist(),
    "simple_exp": rt.construct_t(0, 80, 0.1, 1.0, 0.1, p1, p3, simple_exp=True).tolist(),
    "gaussian_t": rt.construct_t(0, 80, 0.1, 1.0, 0.1, p1, simple_exp=True, gaussian_t=True).tolist(),
}
out["simple_t_gaussian"] = rt.simple_t_gaussian(0, 10, 80, 0.1, 1.0, p1).tolist()
out["round_to_dt"] = rt.round_to_dt(np.array([0.04, 0.11, 0.12, 0.26, 0.31]), 0.1).tolist()
# --- operator strings (tests/test_output_ops.py) and compose_dm
out["output_ops_dm"] = {"2": rt.output_ops_dm(2), "6": rt.output_ops_dm(6), "2_1": rt.output_ops_dm((2, 1)),
                        "2_2": rt.output_ops_dm([2, 2]), "2_2_2": rt.output_ops_dm([2, 2, 2])}
rng = np.random.default_rng(5)
data = rng.standard_normal((7, 4)) + 1j * rng.standard_normal((7, 4))
tt, rho = rt.compose_dm(data, dim=3)
out["compose_dm"] = dict(re=np.real(data).tolist(), im=np.imag(data).tolist(), t=tt.tolist(),
                         rho_re=np.real(rho).tolist(), rho_im=np.imag(rho).tolist())
r = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
r = r @ r.conj().T
r /= np.trace(r)
out["concurrence"] = dict(re=np.real(r).tolist(), im=np.imag(r).tolist(), value=float(rt.concurrence(r)))
bell = np.zeros((4, 4), dtype=complex)
bell[0, 0] = bell[3, 3] = bell[0, 3] = bell[3, 0] = 0.5
out["concurrence_bell"] = float(rt.concurrence(bell))
m = rt.op_to_matrix("(|1><0|_3)")
out["op_to_matrix"] = np.real(m).tolist()

# --- dynamical-map algebra (tools.py:446-675): seeded maps-since-t0 -> time-local maps, pieces, chains
n, nt = 2, 24
times = np.round(0.1 * np.arange(nt), 6)
rng = np.random.default_rng(11)
steps = [np.eye(n * n) + 0.08 * (rng.standard_normal((n * n, n * n)) + 1j * rng.standard_normal((n * n, n * n)))
         for _ in range(nt)]
dm = [steps[0]]
for k in range(1, nt):
    dm.append(steps[k] @ dm[-1])
dm = np.array(dm)
tl = rt.calc_tl_dynmap_pseudo(dm, times)
tl_map, pieces = rt.extract_dms(tl, times, 0.5, [1.0])
rho0 = np.array([[0.6, 0.1 - 0.2j], [0.1 + 0.2j, 0.4]])
cplx = lambda a: dict(re=np.real(a).tolist(), im=np.imag(a).tolist())
out["dynmap"] = dict(times=times.tolist(), dm=cplx(dm), tl=cplx(tl), tl_map=cplx(tl_map),
                     pieces=[cplx(x) for x in pieces], rho0=cplx(rho0),
                     use_tl_map=cplx(rt.use_tl_map(tl_map, times, rho0)),
                     use_dm_block=cplx(rt.use_dm_block(pieces[0], rho0)),
                     use_tl_map_mto=cplx(rt.use_tl_map_mto(tl_map, pieces[0], pieces[1], times, rho0, 1.0)),
                     tl_pad_stationary=cplx(rt.tl_pad_stationary(tl_map, times, rt.use_dm_block(pieces[0], rho0))))

# --- the reference's own parameter / pulse / rotating-frame files (general_system.py:55-102,227-296), written
#     by its prepare_only mode; the temp dir is replaced by the placeholder <TMP>
import tempfile  # noqa: E402
from pyaceqd.four_level_system.linear import biexciton as ref_biexciton  # noqa: E402
from pyaceqd.two_level_system.tls import tls as ref_tls  # noqa: E402

tmp = tempfile.mkdtemp() + "/"
pb = rp.ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
mtos = [{"operator": "|1><3|_4", "applyFrom": "_left", "applyBefore": "false", "time": 2.5},
        {"operator": "|3><1|_4", "applyFrom": "_right", "applyBefore": "false", "time": 2.5},
        {"operator": "|0><1|_4", "applyFrom": "", "applyBefore": "true", "time": 4.0}]
ref_biexciton(0, 8.0, pb, dt=0.25, lindblad=True, delta_b=4.0, delta_xy=0.1, temp_dir=tmp, suffix="k7", multitime_op=mtos,
              output_ops=["|1><1|_4", "|0><3|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)"], prepare_only=True)
pc = rp.ChirpedPulse(tau_0=1.5, e_start=1.0, alpha=3.0, t0=4.0, e0=2.0)
ref_tls(0, 8.0, pc, dt=0.25, lindblad=True, rf=True, temp_dir=tmp, suffix="rf", prepare_only=True)
files = {}
for name in sorted(os.listdir(tmp)):
    files[name] = open(tmp + name).read().replace(tmp, "<TMP>")
out["prepare_only_files"] = files

# --- pulse shaper (pulsegenerator.py): fields built in time / frequency, filters, power, rotating frame, files
import pyaceqd.pulsegenerator as rpg  # noqa: E402


def _pg_state(g):
    return dict(tx=cplx(g.temporal_representation_x), ty=cplx(g.temporal_representation_y),
                fx=cplx(g.frequency_representation_x), fy=cplx(g.frequency_representation_y),
                filt_x=cplx(g.frequency_filter_x), filt_y=cplx(g.frequency_filter_y), power=float(g.pulse_power),
                actions=int(g.action_counter), central_wavelength=float(g.central_wavelength))


pg_out = {}
g = rpg.PulseGenerator(0, 60, 0.25, central_wavelength=800)
pg_out["grid"] = dict(time=g.time.tolist(), frequencies=g.frequencies.tolist(), energies=g.energies.tolist(),
                      wavelengths=g.wavelengths.tolist())
g.add_gaussian_time(width_t=4, central_f=0.1, t0=30, area_time=3.0, sig_or_fwhm='fwhm', field_or_intesity='int',
                    polarisation=[1, 0.5], phase=0.2)
pg_out["gauss_time"] = _pg_state(g)
g.add_filter_double_erf(central_f=0, width_f=0.14, rise_f=0.01)
g.apply_frequency_filter()
pg_out["double_erf_applied"] = _pg_state(g)
g.set_pulse_power(2.5)
pg_out["set_power"] = _pg_state(g)
h = rpg.PulseGenerator(0, 60, 0.25, central_wavelength=800)
h.add_gaussian_freq(width_f=0.5, central_f=-1.0, area_time=2.0, phase_taylor=[0.3, 0.0, 20.0], shift_time=25.0, unit='meV',
                    polarisation=[0, 1])
h.add_filter_gaussian(central_f=-0.2, width_f=0.3, transmission=0.8, sig_fwhm='fwhm', unit='meV', polarisation='y')
h.add_filter_sigmoid(central_f=0.05, width_f=0.2, rise_f=0.01, transmission=0.6, merging='m')
h.add_filter_rectangle(central_f=0.0, width_f=0.6, transmission=0.9, merging='*')
h.add_phase_filter(central_f=0.0, phase_taylor=[0.0, 1.5, -8.0])
pg_out["freq_filters"] = _pg_state(h)
h.apply_frequency_filter('y')
pg_out["freq_filters_applied"] = _pg_state(h)
h.set_rotating_frame(799.0)
pg_out["rotating_frame"] = _pg_state(h)
g.merge_pulses(h)
pg_out["merged"] = _pg_state(g)
pg_tmp = tempfile.mkdtemp() + "/"
fx, fy = g.generate_pulsefiles(temp_dir=pg_tmp, suffix="7")
pg_out["files"] = {os.path.basename(fx): open(fx).read(), os.path.basename(fy): open(fy).read()}
pg_out["units"] = {k: float(g._Units(v, u)) for k, (v, u) in
                   {"meV": (1.3, "meV"), "nm_abs": (801.0, "nm"), "nm_rel": (-0.7, "nm"), "hz": (0.4, "Hz")}.items()}
out["pulsegenerator"] = pg_out

# --- measured-dot calibration files (tools.py:307-344, pulsegenerator.py:66-86, six_level_system/linear.py:33-34)
CALIBRATION = """[EMISSION]
exciton_wavelength = 795.1
biexciton_wavelength = 796.2
dark_wavelength = 795.9
[SPLITTING]
fss_bright = 12.0
fss_dark = 2.0
[LIFETIMES]
exciton = 180
biexciton = 110
dark = 5000
[G_FACTORS]
g_ex = -0.6
g_hx = -0.3
g_ez = -0.75
g_hz = -2.1
"""
cal_tmp = tempfile.mkdtemp() + "/"
with open(cal_tmp + "dot.ini", "w") as fh:
    fh.write(CALIBRATION)
cal = {"text": CALIBRATION, "values": [float(v) for v in rt.read_calibration_file(cal_tmp + "dot.ini")]}
gc = rpg.PulseGenerator(0, 50, 0.5, calibration_file=cal_tmp + "dot.ini")
names = ("central_wavelength", "exciton_x_emission", "exciton_y_emission", "biexciton_x_emission", "biexciton_y_emission",
         "dark_x_emission", "dark_y_emission", "tpe_resonance")
cal["pulsegenerator"] = {n: float(getattr(gc, n)) for n in names}
gc.set_rotating_frame(800.0)
gc.set_rotating_frame(cal_tmp + "dot.ini")
cal["pulsegenerator_after_rf"] = {n: float(getattr(gc, n)) for n in names + ("central_energy", "central_frequency")}
from pyaceqd.six_level_system.linear import sixls_linear as ref_sixls  # noqa: E402
ref_sixls(0, 4.0, pc, dt=0.25, lindblad=True, bx=1.5, bz=0.5, temp_dir=cal_tmp, suffix="cal", prepare_only=True,
          calibration_file=cal_tmp + "dot.ini")
cal["sixls_param"] = {n: open(cal_tmp + n).read().replace(cal_tmp, "<TMP>") for n in sorted(os.listdir(cal_tmp))
                      if n.endswith(".param")}
out["calibration"] = cal

# --- CW spectrum from G1(tau) (two_time/correlations.py:322-382)
import pyaceqd.two_time.correlations as rc  # noqa: E402
tau_s = np.linspace(0, 40.0, 161)
g1_s = (0.3 * np.exp(-0.1 * tau_s) * np.exp(-1j * 0.8 * tau_s) + 0.1 * np.exp(-0.02 * tau_s) + 0.05).astype(complex)
s_om, e_om = rc.get_spectrum(g1_s, tau_s)
out["get_spectrum"] = dict(tau=tau_s.tolist(), g1=cplx(g1_s), s=np.asarray(s_om).tolist(), e=np.asarray(e_om).tolist())

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_host.json"), "w") as fh:
    json.dump(out, fh, default=lambda o: o.item() if hasattr(o, "item") else o.tolist())
print("wrote reference_host.json")
