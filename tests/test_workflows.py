"""Batch issuers: two_time.G1, pol_entanglement.G2, timebin, rabi / tpe sweeps.

CPU tests run the workflows on the oracle backend (host logic: job construction, tail indexing,
integration, density-matrix assembly).  GPU tests run the SAME scenarios on the CUDA engine and
compare every returned array with the oracle-backend result (workflow-level parity, <= 1e-10)."""
import numpy as np
import pytest

from oracle_backend import oracle_backend
from pyaceqd_b200.pulses import ChirpedPulse

TOL = 1e-10
# pseudo-inverses of maps that lost rank to an operator amplify the 1e-13 engine-vs-oracle difference
# the phonon time-local routes divide by dynamical maps (pseudo-inverses with condition numbers ~1e7-1e8): the
# 1e-14 differences between the two backends' propagations come back as ~1e-7 in the (t, tau) grids
WORKFLOW_TOL = {"tl_corr_phonons": 1e-6, "purity_phonons": 1e-6}


# ------------------------------------------------------------------------------------ scenarios
def scenario_g1(tmp):
    from pyaceqd_b200.two_time.G1 import G1_twols
    p = ChirpedPulse(tau_0=1.0, e_start=0.3, alpha=0, t0=3.0, e0=2.0)
    t, tau, g1 = G1_twols(0, 4, 0, 3, 0.5, 0.1, p, gamma_e=0.05, temp_dir=tmp)
    return {"t": t, "tau": tau, "g1": g1}


def _polent(tmp):
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.pol_entanglement.G2 import PolarizatzionEntanglement
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=5.0, polar_x=np.sqrt(0.5))
    opts = {"lindblad": True, "gamma_e": 0.1, "delta_b": 4.0, "delta_xy": 0.2, "phonons": False, "temp_dir": tmp}
    return PolarizatzionEntanglement(biexciton, "|0><1|_4 + |1><3|_4", "|0><2|_4 + |2><3|_4",
                                     "|1><0|_4 + |3><1|_4", "|2><0|_4 + |3><2|_4", p, dt=0.25, tend=10,
                                     regular_grid=True, dt_small=2.5, options=opts)


def scenario_polent(tmp):
    pe = _polent(tmp)
    t1, g2_t, g2 = pe.G2_reuse(pe.axdag, [pe.axdag + " * " + pe.ax, pe.aydag + " * " + pe.ay], pe.ax)
    _, g2_single_t, g2_single = pe.G2(pe.axdag, pe.aydag, pe.ay, pe.ax)
    conc, rho = pe.calc_densitymatrix_reuse(return_rho=True)
    t, c_t, rho_t, norm, rho_int, c_int = pe.calc_timedependent_rho(skip=2)
    _, tau, g1 = pe.G1(pe.ax, pe.axdag)
    return {"t1": t1, "g2_t": g2_t, "g2": g2, "g2_single_t": g2_single_t, "g2_single": g2_single, "conc": conc,
            "rho": rho, "c_t": c_t, "rho_t": rho_t, "c_int": c_int, "g1": g1}


def _timebin(tmp):
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.timebin.twophoton_new import TwoPhotonTimebinNew
    tb = 4.0
    p1 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5, e0=4.0)
    p2 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5 + tb, e0=4.0)
    opts = {"lindblad": True, "gamma_e": 0.5, "delta_b": 4.0, "phonons": False, "temp_dir": tmp}
    return TwoPhotonTimebinNew(biexciton, "|0><1|_4", "|1><0|_4", "|1><3|_4", "|3><1|_4", p1, p2, dt=0.25, dim=4,
                               tb=tb, dt_small=0.5, n_tbig=2, simple_exp=False, options=opts)


def scenario_timebin(tmp):
    tbn = _timebin(tmp)
    conc, rho = tbn.calc_densitymatrix()
    t1, g2, tot, grid = tbn.four_time([tbn.sigma_x, tbn.sigma_x + "*" + tbn.sigma_b],
                                      {"operator": tbn.sigma_bdag, "applyFrom": "_right", "applyBefore": "false"},
                                      {"operator": tbn.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"},
                                      {"operator": tbn.sigma_b, "applyFrom": "_left", "applyBefore": "false"})
    return {"t1": t1, "conc": conc, "rho": rho, "four_time": grid, "g2": g2, "tot": tot}


def scenario_onephoton(tmp):
    from pyaceqd_b200.timebin.onephoton import OnePhotonTimebin
    from pyaceqd_b200.two_level_system.tls import tls
    tb = 4.0
    pulses = [ChirpedPulse(tau_0=0.25, e_start=0, alpha=0, t0=1.5 + k * tb, e0=0.5) for k in range(2)]
    # gaussian_t: the default grid path binds the first pulse to construct_t's dt_exp
    one = OnePhotonTimebin(tls, "|0><1|_2", *pulses, dt=0.1, tb=tb, simple_exp=False, gaussian_t=3.0,
                           options={"lindblad": True, "gamma_e": 0.8, "phonons": False, "temp_dir": tmp})
    ee, ll, el, norm = one.calc_densitymatrix()
    return {"ee": ee, "ll": ll, "el": el, "norm": norm}


def scenario_rabi(tmp):
    from pyaceqd_b200.four_level_system.tpe_rotations import TPERotations
    from pyaceqd_b200.two_level_system.rabi_rotations import RabiRotations
    rr = RabiRotations(dt=0.1, tau=1.0, area_max=4, n_area=9, temp_dir=tmp)
    areas, final = rr.get_rabi_rotations(integrate=False, path=tmp)
    tp = TPERotations(dt=0.25, tau=1.0, area_max=6, n_area=4, gamma_e=0.5)
    a2, xyb = tp.get_rabi_rotations(detuning=-2.0, integrate=True, path=tmp + "i_")
    return {"areas": areas, "final": final, "tpe": xyb}


def _indist(tmp, dm):
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.purity import Indistinguishability
    p = ChirpedPulse(tau_0=0.15, e_start=0, alpha=0, t0=0.8, e0=1.0)
    opts = {"lindblad": True, "gamma_e": 1.5, "phonons": False, "temp_dir": tmp}
    return Indistinguishability(tls, "|0><1|_2", "|1><0|_2", p, dt=0.1, tb=4.0, dt_small=0.1, gaussian_t=2.0,
                                simple_exp=False, dt_big=0.5, options=opts, dm=dm)


def scenario_purity(tmp):
    direct = _indist(tmp, dm=False)
    ind, pur = direct.calc_indistinguishability()
    t2, g2 = direct.G2()
    tl = _indist(tmp, dm=True)
    ind_tl, pur_tl = tl.calc_indistinguishability()
    return {"t1": direct.t1, "ind": ind, "pur": pur, "pur2": direct.calc_purity(), "g2": g2, "ind_tl": ind_tl,
            "pur_tl": pur_tl}


def scenario_dynmap(tmp):
    """Dynamical map (general_system.py:328-335 consumers) applied to a state other than the problem's
    initial one must equal the direct run from that state."""
    from pyaceqd_b200.two_level_system.tls import tls
    p = ChirpedPulse(tau_0=0.5, e_start=0.4, alpha=0, t0=1.5, e0=1.3)
    kw = dict(dt=0.1, lindblad=True, gamma_e=0.3, temp_dir=tmp)
    res, E = tls(0, 4.0, p, calc_dynmap=True, **kw)
    rho0 = np.array([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, 0.7]])
    direct = tls(0, 4.0, p, rho0=rho0, **kw)
    return {"E": E, "via_map": E @ rho0.reshape(-1), "direct": direct}


def scenario_tl_correlations(tmp):
    """Time-local-map G1 / G2 (reference correlations.py:450-864) next to the direct MTO sweeps."""
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.correlations import (three_op_two_time, tl_three_op_two_time, tl_two_op_two_time,
                                                    two_op_two_time)
    # the pulse is over before the earliest run ends: the direct sweeps evaluate the drive of a run's LAST step on
    # that run's own pulse file (end value held, like the reference), which the map route does not imitate
    p = ChirpedPulse(tau_0=0.2, e_start=0.3, alpha=0, t0=0.8, e0=1.7)
    t_axis = np.round(np.arange(0.0, 3.0, 0.5), 6)
    opts = {"lindblad": True, "phonons": False, "gamma_e": 0.4, "temp_dir": tmp}
    rho0 = np.array([[1, 0], [0, 0]], dtype=complex)
    _, tau, g1_tl = tl_two_op_two_time(tls, t_axis, p, tau_max=2.0, dt=0.1, rho0=rho0, options=dict(opts), use_dm=True)
    _, _, g2_tl = tl_three_op_two_time(tls, t_axis, p, tau_max=2.0, dt=0.1, rho0=rho0, options=dict(opts), use_dm=True)
    _, _, g2_f = tl_three_op_two_time(tls, t_axis, p, tau_max=2.0, dt=0.1, rho0=rho0, options=dict(opts), use_dm=True,
                                      fortran_only=True)
    _, _, g1 = two_op_two_time(tls, t_axis, p, tau_max=2.0, dt=0.1, options=dict(opts))
    _, _, g2 = three_op_two_time(tls, t_axis, p, tau_max=2.0, dt=0.1, options=dict(opts))
    # undriven emitter: the stationary map alone is exact
    free = dict(opts)
    _, _, g2_stat = tl_three_op_two_time(tls, t_axis, t_mem=0.5, tau_max=2.0, dt=0.1,
                                         rho0=np.array([[0.2, 0.1], [0.1, 0.8]], dtype=complex), options=free)
    return {"tau": tau, "g1_tl": g1_tl, "g2_tl": g2_tl, "g2_f": g2_f, "g1": g1, "g2": g2, "g2_stat": g2_stat}


PHONON_OPTS = {"lindblad": True, "phonons": True, "t_mem": 1.0, "ae": 5.0, "temperature": 20, "threshold": 6,
               "use_infinite": True}


def scenario_tl_correlations_phonons(tmp):
    """Phonon variants of the time-local G2 (reference correlations.py:866-1185) next to the direct MTO sweep: a
    CW-driven emitter coupled to a QD phonon bath whose process tensor remembers 1 ps."""
    from pyaceqd_b200.pulses import CWLaser
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.correlations import (three_op_two_time, tl_three_op_two_time_phonons,
                                                    tl_threeoptwotime_phonons_dm)
    p = CWLaser(e0=0.6)
    opts = dict(PHONON_OPTS, gamma_e=0.3, temp_dir=tmp, pt_file=tmp + "cw.pt")
    t_axis = np.round(np.arange(0, 4.01, 0.5), 6)
    rho0 = np.array([[1, 0], [0, 0]], dtype=complex)
    _, tau, g2 = three_op_two_time(tls, t_axis, p, tau_max=3.0, dt=0.1, options=dict(opts))
    _, _, g2_tl = tl_three_op_two_time_phonons(tls, t_axis, p, t_mem=1.5, tau_max=3.0, dt=0.1, rho0=rho0,
                                               options=dict(opts))
    _, _, g2_dm = tl_threeoptwotime_phonons_dm(tls, t_axis, p, t_mem=1.5, tau_max=3.0, dt=0.1, rho0=rho0,
                                               options=dict(opts))
    free = dict(opts, phonons=False)
    free.pop("pt_file")
    _, _, g2_free = three_op_two_time(tls, t_axis, p, tau_max=3.0, dt=0.1, options=free)
    return {"tau": tau, "g2": g2, "g2_tl": g2_tl, "g2_dm": g2_dm, "g2_free": g2_free}


def _indist_phonons(tmp, dm):
    from pyaceqd_b200.two_level_system.tls import tls
    from pyaceqd_b200.two_time.purity import Indistinguishability
    p = ChirpedPulse(tau_0=0.15, e_start=0, alpha=0, t0=0.8, e0=1.0)
    opts = dict(PHONON_OPTS, gamma_e=1.5, temp_dir=tmp, pt_file=tmp + "train.pt")
    return Indistinguishability(tls, "|0><1|_2", "|1><0|_2", p, dt=0.05, tb=4.0, dt_small=0.05, gaussian_t=2.5,
                                simple_exp=False, dt_big=0.5, options=opts, dm=dm, t_mem=1.0)


def scenario_purity_phonons(tmp):
    """Indistinguishability with phonons: direct route, the time-local route as the reference's Fortran walks the
    periods (purity.py:513-712) and the same with the period boundary aligned (constants.phonon_block_aligned)."""
    import pyaceqd_b200.constants as constants
    direct = _indist_phonons(tmp, dm=False)
    ind, pur = direct.calc_indistinguishability()
    _, _, grid = direct.G2(return_whole=True)
    tl = _indist_phonons(tmp, dm=True)
    ind_tl, pur_tl = tl.calc_indistinguishability()
    mtos = [{"operator": tl.sigma_x, "applyFrom": "_left", "applyBefore": "false"},
            {"operator": tl.sigma_xdag, "applyFrom": "_right", "applyBefore": "false"}]
    a, c = tl.sigma_xdag_mat, tl.sigma_x_mat
    constants.phonon_block_aligned = True
    try:
        _, grid_al = tl._tl_phonon_grid(mtos, a, a @ c, c)
        ind_al, pur_al = tl.calc_indistinguishability()
    finally:
        constants.phonon_block_aligned = False
    _, rho = tl.calc_timedynamics_tl_phonons()
    occ = direct.calc_timedynamics()[2]
    return {"ind": ind, "pur": pur, "ind_tl": ind_tl, "pur_tl": pur_tl, "ind_al": ind_al, "pur_al": pur_al,
            "grid": grid, "grid_al": np.abs(grid_al), "occ_tl": rho[:, 1, 1].real, "occ": occ.real}


def scenario_timebin_tl(tmp):
    """Time-local route of the time-bin density matrix next to the direct (multi-time-operator) route."""
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.timebin.twophoton_new import TwoPhotonTimebinNew
    tb = 4.0
    p1 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5, e0=4.0)
    p2 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5 + tb, e0=4.0)
    opts = {"lindblad": True, "gamma_e": 0.5, "delta_b": 4.0, "phonons": False, "temp_dir": tmp, "initial": "|0><0|_4"}
    tbn = TwoPhotonTimebinNew(biexciton, "|0><1|_4", "|1><0|_4", "|1><3|_4", "|3><1|_4", p1, p2, dt=0.25, dim=4, tb=tb,
                              dt_small=0.5, n_tbig=2, simple_exp=False, gaussian_t=3.0, simple_t=True, options=opts)
    conc_tl, rho_tl, _ = tbn.calc_densitymatrix_tl(reduced=True)
    _, _, eell_4, grid_4 = tbn.eell_tl_f()
    _, _, eell_8, grid_8 = tbn.eell_tl_8ops()
    _, _, _, d00, g1, g2, _ = tbn.rho_ee_ee(use_second_zero=True)
    _, _, d03, _, _, _ = tbn.rho_ee_ll(use_second_zero=True)
    return {"rho_tl": rho_tl, "conc_tl": conc_tl, "eell_4": eell_4, "eell_8": eell_8, "grid_4": grid_4, "grid_8": grid_8,
            "direct_00": d00, "direct_03": d03}


def scenario_adapters(tmp):
    """Remaining system adapters: sensors / cavities around the two-level emitter, 3- and 4-level dark models."""
    from pyaceqd_b200.four_level_system.dark_model import darkmodel
    from pyaceqd_b200.two_level_system.reduced_dark import darkmodel as darkmodel3
    from pyaceqd_b200.two_level_system.tls import tls, tls_one_sensor, tls_photon, tls_photons, tls_two_sensor
    p = ChirpedPulse(tau_0=0.5, e_start=0, alpha=0, t0=2.0, e0=1.0, polar_x=0.8)
    kw = dict(dt=0.1, lindblad=True, gamma_e=0.2, temp_dir=tmp)
    return {"tls": tls(0, 5, p, **kw)[2],
            "one": tls_one_sensor(0, 5, p, epsilon=1e-6, **kw)[2],
            "two": tls_two_sensor(0, 5, p, epsilon=0.05, linewidth1=0.5, delta_s2=0.3,
                                  output_ops=["|1><1|_2 otimes Id_2 otimes Id_2", "Id_2 otimes |1><1|_2 otimes Id_2",
                                              "Id_2 otimes Id_2 otimes |1><1|_2"], **kw),
            "cav": tls_photon(0, 5, p, n_phot1=2, cav_coupl1=0.4, delta_cx1=0.0, cav_loss1=0.3,
                              output_ops=["|1><1|_2 otimes Id_3", "Id_2 otimes n_3"], **kw),
            "cav2": tls_photons(0, 5, p, n_phot1=1, n_phot2=1, cav_coupl1=0.3, delta_cx1=0, delta_cx2=0.5, **kw),
            "dark4": darkmodel(0, 5, p, delta_xd=0.5, **kw), "dark3": darkmodel3(0, 5, p, delta_xd=0.5, **kw)}


SCENARIOS = {"tl_corr_phonons": scenario_tl_correlations_phonons, "purity_phonons": scenario_purity_phonons, "adapters": scenario_adapters, "timebin_tl": scenario_timebin_tl, "tl_corr": scenario_tl_correlations, "dynmap": scenario_dynmap, "purity": scenario_purity, "g1": scenario_g1, "polent": scenario_polent, "timebin": scenario_timebin,
             "onephoton": scenario_onephoton, "rabi": scenario_rabi}


def _run_oracle(name, tmp_path):
    with oracle_backend() as eng:
        out = SCENARIOS[name](str(tmp_path) + "/o_")
    return out, eng


# ------------------------------------------------------------------------------------ CPU: host logic
def test_g1_layout_and_tau0(tmp_path):
    out, eng = _run_oracle("g1", tmp_path)
    assert out["g1"].shape == (len(out["t"]), 31) and np.allclose(out["tau"], np.linspace(0, 3, 31))
    assert len(eng.calls) == 1 and eng.calls[0][0] == len(out["t"])          # ONE batch for the sweep
    assert np.abs(out["g1"][:, 0].imag).max() < 1e-14 and out["g1"][:, 0].real.min() > -1e-12   # G1(t,0) = x(t)
    # |G1(t,tau)|^2 <= x(t) x(t+tau) (Cauchy-Schwarz); spot check the weaker |G1| <= 1
    assert np.abs(out["g1"]).max() <= 1.0


def test_polent_reuse_equals_single_and_assembly(tmp_path):
    out, eng = _run_oracle("polent", tmp_path)
    assert np.allclose(out["t1"], [0, 2.5, 5.0, 7.5, 10.0])
    # component 1 of the reuse sweep == the stand-alone G2 with the same four operators
    assert np.abs(out["g2_t"][1] - out["g2_single_t"]).max() < 1e-13
    assert abs(out["g2"][1] - out["g2_single"]) < 1e-13
    rho = out["rho"]
    assert np.abs(rho - rho.conj().T).max() < 1e-14 and np.all(np.diag(rho).real >= 0)
    assert 0.0 <= out["conc"] <= 1.0 + 1e-12 and 0.0 <= out["c_int"] <= 1.0 + 1e-12
    assert out["rho_t"].shape == (3, 4, 4) and out["g1"].shape == (5, 41)


def test_integrate_timedep_g2_matches_reference_loops():
    """Vectorised G2(t) = int_0^t dt' int_0^{t-t'} dtau G2(t',tau) against the reference's triple loop
    (pol_entanglement/G2.py:552-606), restated here."""
    from pyaceqd_b200.pol_entanglement.G2 import PolarizatzionEntanglement
    rng = np.random.default_rng(3)
    t1 = np.array([0, 0.5, 1.0, 2.0, 2.5, 4.0])
    t2 = np.linspace(0, 4, 17)
    full = rng.standard_normal((3, len(t1), len(t2))) + 1j * rng.standard_normal((3, len(t1), len(t2)))
    _, got = PolarizatzionEntanglement.integrate_timedep_G2(None, t1, t2, full)
    want = np.zeros((3, len(t1)), dtype=complex)
    for i in range(len(t1)):
        inner = np.zeros((3, i + 1), dtype=complex)
        for j in range(i + 1):
            idx = t2 <= t1[i] - t1[j]
            inner[:, j] = np.trapezoid(full[:, j, idx], t2[idx])
        want[:, i] = np.trapezoid(inner, t1[:i + 1])
    assert np.abs(got - want).max() < 1e-13
    _, g_tau = PolarizatzionEntanglement.integrate_g2_tau(None, t1, t2, full)
    assert np.abs(g_tau[1, 4] - np.trapezoid(full[1, :, 4], t1)) < 1e-14


def test_symmetrised_spectrum_matches_loops():
    from pyaceqd_b200.sweeps import symmetrised_spectrum
    rng = np.random.default_rng(5)
    t, tau = np.array([0, 1.0, 1.5, 3.0]), np.linspace(0, 2, 9)
    g1 = rng.standard_normal((4, 9)) + 1j * rng.standard_normal((4, 9))
    e, spec, spectra = symmetrised_spectrum(t, tau, g1, 0.6582119569)
    sym = np.empty((4, 17), dtype=complex)
    sym[:, :9] = g1[:, ::-1]
    sym[:, -8:] = np.conj(g1[:, 1:])
    want = np.array([np.fft.fftshift(np.fft.fft(sym[j])) for j in range(4)])
    assert np.abs(spectra - want).max() < 1e-12
    assert np.abs(spec - np.real(np.trapezoid(want.T, t))).max() < 1e-12
    assert len(e) == 17 and np.all(np.diff(e) < 0)      # energies = -2 pi hbar f, shifted


def test_timebin_density_matrix_structure(tmp_path):
    out, eng = _run_oracle("timebin", tmp_path)
    rho, n = out["rho"], len(out["t1"])
    assert np.abs(rho - rho.conj().T).max() < 1e-14
    assert np.all(np.diag(rho).real > 0) and 0.0 <= out["conc"] <= 1.0 + 1e-12
    # identical pulses in both bins -> similar early/late populations (decay is not complete within tb here)
    assert abs(rho[0, 0] - rho[3, 3]) < 0.25 * abs(rho[0, 0])
    # triangular sweeps are issued as few large batches, not one per t1
    assert max(c[0] for c in eng.calls) == n * (n + 1) // 2
    assert np.allclose(np.tril(out["four_time"], -1), 0)


def test_onephoton_and_area_sweeps(tmp_path):
    out, _ = _run_oracle("onephoton", tmp_path)
    assert 0 < out["ee"] < out["ll"] < 2 * out["ee"]      # late bin also collects the early bin's leftover
    assert 0 < out["el"] <= np.sqrt(out["ee"] * out["ll"]) + 1e-9    # |rho_el| <= sqrt(rho_ee rho_ll) (gamma_e units)
    out, eng = _run_oracle("rabi", tmp_path)
    assert np.abs(out["final"] - np.sin(np.pi * out["areas"] / 2) ** 2).max() < 2e-3    # Rabi rotations
    assert out["tpe"].shape == (3, 4) and np.all(out["tpe"] >= -1e-12)
    assert [c[0] for c in eng.calls] == [9, 4]


def test_dynmap_acts_on_arbitrary_states(tmp_path):
    out, _ = _run_oracle("dynmap", tmp_path)
    assert out["E"].shape == (40, 4, 4)
    # outputs of tls: g = rho_00, x = rho_11, <|0><1|> = rho_10, <|1><0|> = rho_01 (row-major vec: 00, 01, 10, 11)
    assert np.abs(out["via_map"][:, 0] - out["direct"][1][1:]).max() < 1e-12
    assert np.abs(out["via_map"][:, 3] - out["direct"][2][1:]).max() < 1e-12
    assert np.abs(out["via_map"][:, 2] - out["direct"][3][1:]).max() < 1e-12
    assert np.linalg.matrix_rank(out["E"][5]) == 4


def test_composite_and_dark_adapters(tmp_path):
    out, _ = _run_oracle("adapters", tmp_path)
    assert np.abs(out["one"] - out["tls"]).max() < 1e-9            # a sensor coupled with epsilon -> 0 does not act back
    x, s1, s2 = out["two"][1].real, out["two"][2].real, out["two"][3].real
    assert s1.max() > 1e-4 and s2.max() > 1e-5 and s1.max() < 0.05 * x.max()       # sensors pick up a little signal
    assert out["cav"][2].real.max() > 0.05 and np.abs(out["cav"][1].real).max() <= 1 + 1e-9   # photons appear
    assert out["cav2"].shape[0] == 3
    for key, n in (("dark4", 4), ("dark3", 3)):
        pops = out[key][1:1 + n].real
        assert np.abs(pops.sum(axis=0) - 1).max() < 1e-10 and pops.min() > -1e-12          # trace, positivity
    from pyaceqd_b200.two_level_system.tls import tls_photon_sensor, tls_photon_two_sensor
    with pytest.raises(ValueError):
        tls_photon_sensor(0, 1, ChirpedPulse(tau_0=0.5, e_start=0, alpha=0, t0=2.0, e0=1.0), n_phot1=2)
    with pytest.raises(ValueError):
        tls_photon_two_sensor(0, 1)


def test_timebin_time_local_route(tmp_path):
    out, _ = _run_oracle("timebin_tl", tmp_path)
    rho = out["rho_tl"]
    assert np.abs(rho - rho.conj().T).max() < 1e-14 and np.all(np.diag(rho).real > 0)
    assert 0.0 <= out["conc_tl"] <= 1.0 + 1e-12
    # four-operator and eight-operator chain programs describe the same element
    assert np.abs(out["grid_4"] - out["grid_8"]).max() < 1e-12 and abs(out["eell_4"] - out["eell_8"]) < 1e-13
    # and agree with the direct multi-time-operator route (first time ordering) up to the quadrature:
    # the direct route integrates t2 on the simulation grid, the map route on the coarse t1 grid
    assert abs(rho[0, 0] - out["direct_00"]) < 5e-2 * abs(out["direct_00"])
    assert abs(rho[0, 3] - out["direct_03"]) < 5e-2 * abs(out["direct_03"]) + 1e-6


def test_tl_correlations_equal_direct_sweeps(tmp_path):
    """Without phonons the quantum regression theorem is exact: chains of time-local maps reproduce the
    multi-time-operator trajectories (up to the pseudo-inverse's conditioning)."""
    out, _ = _run_oracle("tl_corr", tmp_path)
    assert np.abs(out["g1_tl"] - out["g1"]).max() < 1e-8
    assert np.abs(out["g2_tl"] - out["g2"]).max() < 1e-8
    assert np.abs(out["g2_f"] - out["g2"]).max() < 1e-8      # real operators: the column-major reading agrees
    g = out["g2_stat"]
    assert np.allclose(g[:, 0], 0) and np.all(np.abs(g[:, 1:]) < 1e-12)    # G2 of a single emitter stays 0 undriven


def test_tl_correlations_with_phonons_follow_direct_sweeps(tmp_path):
    """With a bath memory the time-local routes are exact up to the memory cut (1.5 ps here for a PT that remembers
    1 ps): they follow the direct sweeps far closer than the phonon-free dynamics do."""
    out, eng = _run_oracle("tl_corr_phonons", tmp_path)
    assert np.abs(out["g2_tl"] - out["g2"]).max() < 5e-5
    assert np.abs(out["g2_dm"] - out["g2"]).max() < 5e-5
    assert np.abs(out["g2_free"] - out["g2"]).max() > 5e-3            # the bath matters in this scenario
    # the dynamical-map runs inside the memory time are ONE batch of runs plus one of their NL unit vectors
    assert any(c[0] == 3 * 4 for c in eng.calls)


def test_indistinguishability_with_phonons(tmp_path):
    out, eng = _run_oracle("purity_phonons", tmp_path)
    # per-period maps reproduce the pulse-train dynamics
    n = min(len(out["occ"]), len(out["occ_tl"]))
    assert np.abs(out["occ"][:n] - out["occ_tl"][:n]).max() < 2e-3
    # aligned at the period boundary, the map route fills the same (t, tau) grid as the direct one (the last t sits
    # ON the boundary and is left out); the reference's own walk returns to the period start one step early
    assert np.abs(out["grid"][:-1] - out["grid_al"][:-1]).max() < 1e-3
    assert abs(out["pur_al"] - out["pur"]) < 2e-3 and abs(out["ind_al"] - out["ind"]) < 2e-3
    assert abs(out["pur_tl"] - out["pur"]) < 1e-2 and abs(out["ind_tl"] - out["ind"]) < 1e-2
    assert any(c[0] == 17 * 4 for c in eng.calls)                     # all t inside the memory: one batch of maps


def test_purity_and_indistinguishability_routes(tmp_path):
    out, eng = _run_oracle("purity", tmp_path)
    assert abs(out["pur"] - out["pur2"]) < 1e-13
    assert 0.5 < out["pur"] <= 1.0 + 1e-9 and 0.3 < out["ind"] <= 1.0 + 1e-9      # short pi-pulse: nearly pure photons
    # the time-local route (per-period dynamical map + chain kernel) reproduces the direct route; both
    # integrate the same (t, tau) grid, the map route cuts the pulse tail at gaussian_t
    assert abs(out["pur_tl"] - out["pur"]) < 5e-3 and abs(out["ind_tl"] - out["ind"]) < 5e-3


def test_planner_tail_rows_and_fork_bookkeeping():
    """Engine.plan with tail_rows: rows kept, out_from, trunk copies (no GPU needed)."""
    from helpers import make_tables, tls_problem
    from pyaceqd_b200.engine import Engine
    from pyaceqd_b200.jobs import Job
    from pyaceqd_b200.process_tensor import trivial_pt
    prob = tls_problem(phonons=False)
    p = ChirpedPulse(tau_0=1.0, e_start=0, alpha=0, t0=2.0, e0=1.0)
    tabs = make_tables([p], 0.0, 6.0, 0.1)
    mto = lambda t: prob.parse_mtos([{"operator": "|0><1|_2", "applyFrom": "_left", "time": t}])
    jobs = [Job(0.0, 4.0, 0.1, tables=tabs, mtos=mto(1.0), tail_rows=11),     # tail entirely on the branch
            Job(0.0, 4.0, 0.1, tables=tabs, mtos=mto(3.5), tail_rows=11),     # tail reaches back into the trunk
            Job(0.0, 5.0, 0.1, tables=tabs, mtos=mto(2.0))]                   # all rows
    eng = Engine.__new__(Engine)
    common, plan = Engine.plan(eng, prob, trivial_pt(1), jobs)
    trunk, main = plan.levels
    assert plan.n_rows.tolist() == [11, 11, 51] and plan.out_elems == (11 + 11 + 51) * prob.n_out
    assert main.job.tolist() == [0, 1, 2]
    assert (main.step0[0], main.out_from[0], main.row0[0]) == (10, 20, 0)     # rows 30..40, branch starts at 10
    assert (main.step0[1], main.out_from[1], main.row0[1]) == (35, 0, 5)      # rows 30..34 from the trunk
    assert (main.step0[2], main.out_from[2], main.row0[2]) == (20, 0, 20)
    assert sorted((c[0], c[1], c[3]) for c in plan.copies.tolist()) == [(1, 5, 30), (2, 20, 0)]
    assert trunk.snap_steps.tolist() == [10, 20, 35] and trunk.n_steps.tolist() == [35]


# ------------------------------------------------------------------------------------ GPU: workflow parity
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_workflow_gpu_equals_oracle_backend(name, tmp_path, monkeypatch):
    # both legs propagate the SAME process tensor (the NumPy builder's): this test is about the workflows and the
    # kernels; the device PT builder has its own parity tests (tests/test_ptbuild.py), and the time-local routes
    # amplify truncation-level differences between two builds of a PT through their pseudo-inverses
    monkeypatch.setenv("ACEQD_PT_BUILD", "host")
    want, _ = _run_oracle(name, tmp_path)
    got = SCENARIOS[name](str(tmp_path) + "/g_")
    assert sorted(got) == sorted(want)
    for k in want:
        d = np.abs(np.asarray(got[k]) - np.asarray(want[k])).max()
        assert d < WORKFLOW_TOL.get(name, TOL), (name, k, d)


@pytest.mark.gpu
def test_tail_rows_equal_tail_of_full_run():
    from helpers import make_tables, tls_problem
    from pyaceqd_b200.engine import default_engine
    from pyaceqd_b200.jobs import Job
    from pyaceqd_b200.process_tensor import synthetic_pt
    prob = tls_problem()
    pt = synthetic_pt(24, len(prob.cls_keys), kind="unitary", scale=0.999)
    p = ChirpedPulse(tau_0=1.0, e_start=0.5, alpha=0, t0=2.0, e0=3.0)
    tabs = make_tables([p], 0.0, 6.0, 0.1)
    mto = lambda t: prob.parse_mtos([{"operator": "|0><1|_2", "applyFrom": "_left", "time": t}])
    mk = lambda tail: [Job(0.0, 4.0, 0.1, tables=tabs, mtos=mto(1.0), tail_rows=tail),
                       Job(0.0, 4.0, 0.1, tables=tabs, mtos=mto(3.5), tail_rows=tail),
                       Job(0.0, 5.0, 0.1, tables=tabs, mtos=mto(2.0), tail_rows=tail),
                       Job(0.0, 3.0, 0.1, tables=tabs, tail_rows=tail)]
    eng = default_engine(0)
    full = eng.run_jobs(prob, pt, mk(0))
    for kernel in ("dmma", "check"):
        for fork in (True, False):
            tails = eng.run_jobs(prob, pt, mk(11), kernel=kernel, fork=fork)
            for f, t in zip(full, tails):
                assert t.shape[1] == 11 and np.abs(f[:, -11:] - t).max() < 1e-12


def test_timebin_four_time_tl_kernel_route_equals_the_python_route(tmp_path):
    """``four_time_tl`` (reference ``timebin/twophoton_new.py:925-1013``: a double Python loop over (t1, t2) of
    ``propagate_tb_new`` calls with operator insertions and a trace) against the chain-kernel route this package
    gives the same method, on the oracle backend; plus the time-local dynamics helpers."""
    from pyaceqd_b200.four_level_system.linear import biexciton
    from pyaceqd_b200.timebin.twophoton_new import TwoPhotonTimebinNew
    from pyaceqd_b200.tools import op_to_matrix
    tb = 4.0
    p1 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5, e0=4.0)
    p2 = ChirpedPulse(tau_0=0.25, e_start=-2.0, alpha=0, t0=1.5 + tb, e0=4.0)
    opts = {"lindblad": True, "gamma_e": 0.5, "delta_b": 4.0, "phonons": False, "temp_dir": str(tmp_path),
            "initial": "|0><0|_4"}
    with oracle_backend():
        tbn = TwoPhotonTimebinNew(biexciton, "|0><1|_4", "|1><0|_4", "|1><3|_4", "|3><1|_4", p1, p2, dt=0.25, dim=4,
                                  tb=tb, dt_small=0.5, n_tbig=2, simple_exp=False, gaussian_t=3.0, simple_t=True,
                                  options=opts)
        t1, g2, eell, grid = tbn.four_time_tl(tbn.sigma_bdag, tbn.sigma_xdag, tbn.sigma_b, tbn.sigma_x)
        _, _, eell_f, grid_f = tbn.eell_tl_f()
        assert np.abs(grid - grid_f).max() < 1e-13 and abs(eell - eell_f) < 1e-13
        assert tbn.eell_tl()[2] == eell
        # the reference's Python loop, literally
        s1, s2, s3, s4 = (op_to_matrix(o) for o in (tbn.sigma_bdag, tbn.sigma_xdag, tbn.sigma_b, tbn.sigma_x))
        rho0 = tbn.get_initial_state()
        dim = rho0.shape[0]
        ref = np.zeros((len(t1), len(t1)), dtype=complex)
        for i, a in enumerate(t1):
            r = tbn.propagate_tb_new(0, a, rho0.copy().reshape(dim * dim), tbn.dm_tl1).reshape(dim, dim) @ s1
            for j in range(i, len(t1)):
                b = t1[j]
                q = tbn.propagate_tb_new(a, b, r.copy().reshape(dim * dim), tbn.dm_tl1).reshape(dim, dim) @ s2
                q = tbn.propagate_tb_new(b, tbn.tb, q.reshape(dim * dim), tbn.dm_tl1)
                q = tbn.propagate_tb_new(0, a, q, tbn.dm_tl2).reshape(dim, dim)
                q = s3 @ q
                q = tbn.propagate_tb_new(a, b, q.reshape(dim * dim), tbn.dm_tl2).reshape(dim, dim)
                ref[i, j] = np.trace(s4 @ q)
        assert np.abs(ref).max() > 1e-3
        assert np.abs(grid - ref).max() < 1e-12
        # dynamics helpers: the t1-grid dynamics are the fine-grid dynamics at the t1 points, with trace 1
        t_fine, rho_fine = tbn.dynamics_tl()
        t_c, rho_c = tbn.dynamics_tl_t1()
        assert len(t_c) == 2 * len(t1) - 1 and np.allclose(np.trace(rho_c, axis1=1, axis2=2), 1.0, atol=1e-10)
        for k in (1, len(t1) - 1, len(t1) + 1):
            i = int(round(t_c[k] / tbn.dt))
            if i < len(t_fine):
                assert np.abs(rho_c[k] - rho_fine[i]).max() < 1e-10
        t_i, rho_i = tbn.dynamics_tl_t1_t2_f(t1[1], t1[2], None, None, None, take_IDs=True)
        assert np.abs(rho_i - rho_c).max() < 1e-10
