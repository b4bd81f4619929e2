"""Golden fixtures of the reference's BATCH ISSUERS, made by running the UNMODIFIED reference workflow code
(/root/reference/pyaceqd: two_time.correlations, pol_entanglement.G2, timebin.twophoton_new) in this container.
The reference shells out to `ACE <param file>`; here `ACE` on $PATH is tests/golden/ace_oracle.py, i.e. the
reference-format file reader/writer of this repo with the CPU oracle as propagation backend.  So everything the
reference's host code decides -- which trajectories are issued, where the multi-time operators sit, which output
rows are picked from the END of each run (two_time/correlations.py:182-183), n_t2 = n_tau - int(t1/dt)
(pol_entanglement/G2.py:283), the triangular (t1, t2) bookkeeping (timebin/twophoton_new.py:515-557), the tau/t
integrals -- is executed by the reference itself, and tests/test_reference_workflows.py checks that pyaceqd_b200's
workflows reproduce the stored arrays.

    python tests/golden/make_reference_workflows.py        # writes tests/golden/reference_workflows.npz

matplotlib, ACEutils and the f2py module timebin_tl are absent here and are stubbed in sys.modules (SURVEY App. E
R7); nothing of the reference is edited.  /root/reference does not exist on the GPU box: only the .npz travels.
"""
import os
import stat
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.collections", "ACEutils",
             "pyaceqd.timebin.timebin_tl"):
    sys.modules.setdefault(name, types.ModuleType(name))
for name in ("Parameters", "FreePropagator", "ProcessTensors", "InitialState", "OutputPrinter", "TimeGrid", "Simulation",
             "read_outfile", "DynamicalMap"):
    setattr(sys.modules["ACEutils"], name, None)
import pyaceqd.timebin as _tb  # noqa: E402
_tb.timebin_tl = sys.modules["pyaceqd.timebin.timebin_tl"]

# `ACE` on $PATH -> the oracle-backed executable
bindir = tempfile.mkdtemp(prefix="aceqd_fake_ace_")
with open(os.path.join(bindir, "ACE"), "w") as fh:
    fh.write("#!/bin/sh\nexec {} {} \"$@\"\n".format(sys.executable, os.path.join(HERE, "ace_oracle.py")))
os.chmod(os.path.join(bindir, "ACE"), os.stat(os.path.join(bindir, "ACE")).st_mode | stat.S_IEXEC)
os.environ["PATH"] = bindir + os.pathsep + os.environ["PATH"]

import pyaceqd.pulses as rp  # noqa: E402
from pyaceqd.four_level_system.linear import biexciton as ref_biexciton  # noqa: E402
from pyaceqd.four_level_system.dark_model import darkmodel_new as ref_darkmodel  # noqa: E402
from pyaceqd.pol_entanglement.G2 import PolarizatzionEntanglement  # noqa: E402
from pyaceqd.timebin.twophoton_new import TwoPhotonTimebinNew  # noqa: E402
from pyaceqd.two_level_system.tls import tls as ref_tls  # noqa: E402
from pyaceqd.two_time.correlations import three_op_two_time, two_op_two_time  # noqa: E402

tmp = tempfile.mkdtemp(prefix="aceqd_ref_") + "/"
out = {}

# ---- G1(t, tau) of a driven two-level system: two_time/correlations.py:186-225 -> _ops_two_time :135-184
p = rp.ChirpedPulse(tau_0=0.8, e_start=0.3, alpha=0, t0=2.0, e0=1.5)
t_axis = np.round(np.arange(0.0, 3.0, 0.5), 6)
t1, tau, G = two_op_two_time(ref_tls, t_axis, p, opA="|1><0|_2", opB="|0><1|_2", tau_max=2.0, dt=0.1, workers=4,
                             options={"lindblad": True, "phonons": False, "gamma_e": 0.2, "temp_dir": tmp})
out["g1_t1"], out["g1_tau"], out["g1_G"] = t1, tau, G

# ---- G2(t, tau) of the biexciton: three_op_two_time :227-270
pb = rp.ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
t_axis = np.round(np.arange(0.0, 5.0, 0.75), 6)
t1, tau, G = three_op_two_time(ref_biexciton, t_axis, pb, opA="|3><1|_4", opB="|1><1|_4", opC="|1><3|_4", tau_max=3.0,
                               dt=0.25, workers=4,
                               options={"lindblad": True, "phonons": False, "delta_b": 4.0, "delta_xy": 0.1,
                                        "gamma_e": 0.05, "gamma_b": 0.07, "temp_dir": tmp})
out["g2_t1"], out["g2_tau"], out["g2_G"] = t1, tau, G

# ---- polarisation entanglement: G2_reuse (pol_entanglement/G2.py:439-533) and the reused density matrix (:301-357)
pb = rp.ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=5.0, e0=4.0, polar_x=0.8)   # t grid starts at t0 - 4 tau >= 0
pe = PolarizatzionEntanglement(ref_biexciton, "|0><1|_4 + |1><3|_4", "|0><2|_4 + |2><3|_4", "|1><0|_4 + |3><1|_4",
                               "|2><0|_4 + |3><2|_4", pb, dt=0.25, tend=12.0, simple_exp=True, dt_small=1.0, workers=4,
                               options={"lindblad": True, "phonons": False, "delta_b": 4.0, "delta_xy": 0.1,
                                        "gamma_e": 0.05, "gamma_b": 0.07, "temp_dir": tmp})
out["pe_t1"] = np.asarray(pe.t1, dtype=float)
t1r, t2r, g2r, g2int, g2full = pe.G2_reuse(pe.axdag, [pe.axdag + "*" + pe.ax, pe.aydag + "*" + pe.ay, pe.axdag + "*" + pe.ay],
                                           pe.ax, return_full_G2=True)
out["pe_reuse_t1"], out["pe_reuse_t2"], out["pe_reuse_G2"], out["pe_reuse_int"], out["pe_reuse_full"] = t1r, t2r, g2r, g2int, g2full
conc, rho = pe.calc_densitymatrix_reuse(return_rho=True)
out["pe_concurrence"], out["pe_rho"] = np.asarray(conc), np.asarray(rho)

# ---- time-bin four-time correlation: triangular (t1, t2) sweep with three multi-time operators per run
#      (timebin/twophoton_new.py:515-557)
pt_ = rp.ChirpedPulse(tau_0=0.5, e_start=-2.0, alpha=0, t0=2.0, e0=5.0, polar_x=1.0)   # t grid starts at t0 - 4 tau = 0
tbn = TwoPhotonTimebinNew(ref_darkmodel, "|0><1|_5", "|1><0|_5", "|1><4|_5", "|4><1|_5", pt_, dt=0.1, dim=5, tb=5.0,
                          dt_small=1.0, n_tbig=1, simple_exp=False, workers=4,
                          options={"lindblad": True, "phonons": False, "delta_b": 4.0, "gamma_e": 0.2, "temp_dir": tmp})
out["tb_t1"] = np.asarray(tbn.t1, dtype=float)
# rho_ee_ll (:368-393) = two four_time sweeps with the operator patterns the reference uses itself
rt1, rG, reell, rG1, rG2, rG12 = tbn.rho_ee_ll()
out["tb_eell_t1"], out["tb_eell_G"], out["tb_eell"], out["tb_eell_G1"], out["tb_eell_G2"], out["tb_eell_G12"] = \
    rt1, rG, np.asarray(reell), rG1, rG2, rG12

np.savez_compressed(os.path.join(HERE, "reference_workflows.npz"), **out)
for k, v in out.items():
    print(k, np.asarray(v).shape, float(np.abs(np.asarray(v)).max()))
print("wrote reference_workflows.npz")
