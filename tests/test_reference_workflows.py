"""The batch issuers pinned to the REFERENCE's own execution.

tests/golden/reference_workflows.npz was produced by running the unmodified reference workflow code
(/root/reference/pyaceqd two_time.correlations, pol_entanglement.G2, timebin.twophoton_new) in the build container,
with `ACE` on $PATH resolved to this repo's reference-format reader + the CPU oracle
(tests/golden/make_reference_workflows.py).  What the reference's host code decides -- which trajectories are issued,
multi-time-operator placement, the rows picked from the END of every run (two_time/correlations.py:182-183),
n_t2 = n_tau - int(t1/dt) (pol_entanglement/G2.py:283), the triangular (t1, t2) sweep (timebin/twophoton_new.py:
515-557), the integrals -- is therefore the reference's, and pyaceqd_b200's workflows must reproduce the arrays:
on the CPU with the oracle as engine (host logic alone) and on the GPU through the CUDA path.

Tolerance 2e-10: the reference path goes through text files with 12 significant digits (set_precision 12)."""
import os

import numpy as np
import pytest

from oracle_backend import oracle_backend
from pyaceqd_b200.four_level_system.dark_model import darkmodel_new
from pyaceqd_b200.four_level_system.linear import biexciton
from pyaceqd_b200.pol_entanglement.G2 import PolarizatzionEntanglement
from pyaceqd_b200.pulses import ChirpedPulse
from pyaceqd_b200.timebin.twophoton_new import TwoPhotonTimebinNew
from pyaceqd_b200.two_level_system.tls import tls
from pyaceqd_b200.two_time.correlations import three_op_two_time, two_op_two_time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = np.load(os.path.join(HERE, "golden", "reference_workflows.npz"))
TOL = 2e-10


def _close(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    dev = float(np.abs(a - b).max()) if a.size else 0.0
    assert dev < TOL * max(1.0, float(np.abs(b).max())), f"{what}: max abs deviation {dev:.3e}"


def _g1(tmp):
    p = ChirpedPulse(tau_0=0.8, e_start=0.3, alpha=0, t0=2.0, e0=1.5)
    t_axis = np.round(np.arange(0.0, 3.0, 0.5), 6)
    t1, tau, G = two_op_two_time(tls, t_axis, p, opA="|1><0|_2", opB="|0><1|_2", tau_max=2.0, dt=0.1, workers=4,
                                 options={"lindblad": True, "phonons": False, "gamma_e": 0.2, "temp_dir": tmp})
    _close(t1, REF["g1_t1"], "G1 t axis")
    _close(tau, REF["g1_tau"], "G1 tau axis")
    _close(G, REF["g1_G"], "G1(t, tau)")
    assert np.abs(REF["g1_G"]).max() > 0.1


def _g2(tmp):
    pb = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
    t_axis = np.round(np.arange(0.0, 5.0, 0.75), 6)
    t1, tau, G = three_op_two_time(biexciton, t_axis, pb, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4", tau_max=3.0,
                                   dt=0.25, workers=4,
                                   options={"lindblad": True, "phonons": False, "delta_b": 4.0, "delta_xy": 0.1,
                                            "gamma_e": 0.05, "gamma_b": 0.07, "temp_dir": tmp})
    _close(tau, REF["g2_tau"], "G2 tau axis")
    _close(G, REF["g2_G"], "G2(t, tau)")
    assert np.abs(REF["g2_G"]).max() > 0.1


def _polent(tmp):
    pb = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=5.0, e0=4.0, polar_x=0.8)
    pe = PolarizatzionEntanglement(biexciton, "|0><1|_4 + |1><3|_4", "|0><2|_4 + |2><3|_4", "|1><0|_4 + |3><1|_4",
                                   "|2><0|_4 + |3><2|_4", pb, dt=0.25, tend=12.0, simple_exp=True, dt_small=1.0, workers=4,
                                   options={"lindblad": True, "phonons": False, "delta_b": 4.0, "delta_xy": 0.1,
                                            "gamma_e": 0.05, "gamma_b": 0.07, "temp_dir": tmp})
    _close(pe.t1, REF["pe_t1"], "pol-entanglement t grid")
    t1r, t2r, g2r, g2int, g2full = pe.G2_reuse(pe.axdag, [pe.axdag + "*" + pe.ax, pe.aydag + "*" + pe.ay,
                                                          pe.axdag + "*" + pe.ay], pe.ax, return_full_G2=True)
    _close(t2r, REF["pe_reuse_t2"], "G2_reuse tau axis")
    _close(g2full, REF["pe_reuse_full"], "G2_reuse full G2(t, tau)")
    _close(g2r, REF["pe_reuse_G2"], "G2_reuse tau integrals")
    _close(g2int, REF["pe_reuse_int"], "G2_reuse t and tau integrals")
    conc, rho = pe.calc_densitymatrix_reuse(return_rho=True)
    _close(rho, REF["pe_rho"], "two-photon density matrix")
    _close(conc, REF["pe_concurrence"], "concurrence")
    assert float(REF["pe_concurrence"]) > 0.05


def _timebin(tmp):
    p = ChirpedPulse(tau_0=0.5, e_start=-2.0, alpha=0, t0=2.0, e0=5.0, polar_x=1.0)
    tbn = TwoPhotonTimebinNew(darkmodel_new, "|0><1|_5", "|1><0|_5", "|1><4|_5", "|4><1|_5", p, dt=0.1, dim=5, tb=5.0,
                              dt_small=1.0, n_tbig=1, simple_exp=False, workers=4,
                              options={"lindblad": True, "phonons": False, "delta_b": 4.0, "gamma_e": 0.2, "temp_dir": tmp})
    _close(tbn.t1, REF["tb_t1"], "timebin t grid")
    t1, G, eell, G1, G2, G12 = tbn.rho_ee_ll()
    _close(G12, REF["tb_eell_G12"], "four-time grid G(t1, t2)")
    _close(G1, REF["tb_eell_G1"], "four-time t2 integrals, t1 <= t2")
    _close(G2, REF["tb_eell_G2"], "four-time t2 integrals, t2 <= t1")
    _close(eell, REF["tb_eell"], "<ee|rho|ll>")
    assert np.abs(REF["tb_eell_G12"]).max() > 0.05


CASES = {"g1_two_op_two_time": _g1, "g2_three_op_two_time": _g2, "pol_entanglement_G2_reuse": _polent,
         "timebin_rho_ee_ll_four_time": _timebin}


@pytest.mark.parametrize("name", sorted(CASES))
def test_workflow_host_logic_reproduces_reference_execution(name, tmp_path):
    with oracle_backend():
        CASES[name](str(tmp_path) + "/")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_workflow_on_gpu_reproduces_reference_execution(name, tmp_path):
    CASES[name](str(tmp_path) + "/")
