"""The process boundary (SURVEY 8b, B0 / known-answer test K7): parameter files written exactly like
the reference's (general_system.py:227-290) are parsed back into the same numeric problem, and the
`ACE <file>` executable reproduces the in-process result to the text precision."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_backend import oracle_backend
from pyaceqd_b200 import ace_cli
from pyaceqd_b200.general_system import general_system as gs
from pyaceqd_b200.pulses import ChirpedPulse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _prepare(tmp, **extra):
    from pyaceqd_b200.four_level_system.linear import biexciton
    p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=3.0, e0=4.0, polar_x=0.8)
    mtos = [{"operator": "|1><3|_4", "applyFrom": "_left", "applyBefore": "false", "time": 2.5},
            {"operator": "|3><1|_4", "applyFrom": "_right", "applyBefore": "false", "time": 2.5},
            {"operator": "|0><1|_4", "applyFrom": "", "applyBefore": "true", "time": 4.0}]
    kw = dict(dt=0.25, lindblad=True, delta_b=4.0, delta_xy=0.1, temp_dir=tmp, suffix="k7", multitime_op=mtos,
              output_ops=["|1><1|_4", "|0><3|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)"])
    kw.update(extra)
    biexciton(0, 8.0, p, prepare_only=True, **kw)
    return biexciton, p, kw, tmp + "b_linear_k7.param"


def test_param_file_keys_and_parse_round_trip(tmp_path):
    tmp = str(tmp_path) + "/"
    system, p, kw, param = _prepare(tmp)
    text = open(param).read().splitlines()
    assert text[:6] == ["dt    0.25", "ta    0", "te    8.0", "dict_zero 1e-16", "set_precision 12",
                        "use_symmetric_Trotter true"]
    assert text[6] == "initial    { |0><0|_4 }" and text[-1] == "outfile " + tmp + "b_linear_k7.out"
    assert "apply_Operator_left 2.5 { |1><3|_4 } false" in text and "apply_Operator 4.0 { |0><1|_4 } true" in text
    assert any(l.startswith("add_Pulse file " + tmp + "b_linear_pulse_y_k7.dat  { -0.5*pi*hbar*(|2><0|_4+|3><2|_4) }")
               for l in text)
    params = ace_cli.parse_param_file(param)
    prob, tables, mtos = ace_cli.problem_from_params(params)
    direct = gs._problem_for(system_op=["-4.0*|3><3|_4", "-0.05*|1><1|_4", "0.05*|2><2|_4"], boson_op=None,
                             initial="|0><0|_4",
                             lindblad_ops=[["|0><1|_4", 0.01], ["|0><2|_4", 0.01], ["|1><3|_4", 0.01], ["|2><3|_4", 0.01]],
                             interaction_ops=[["|1><0|_4+|3><1|_4", "x"], ["|2><0|_4+|3><2|_4", "y"]],
                             output_ops=kw["output_ops"], rf_op=None, rho0=None, dict_zero="16")
    for a in ("L0", "LA", "LB", "rho0", "out_w"):
        assert np.abs(getattr(prob, a) - getattr(direct, a)).max() < 1e-14, a
    assert prob.field_pol == ["x", "y"] and [m["applyFrom"] for m in mtos] == ["_left", "_right", ""]
    # pulse file = the %.8f samples on np.arange(t_start, t_end, dt)
    t = np.arange(0, 8.0, 0.25)
    px, py = gs.sample_pulses(t, [p])
    assert np.abs(tables["x"].values - px).max() < 1e-15 and np.abs(tables["y"].values - py).max() < 1e-15
    assert tables["x"].t0 == 0.0 and tables["x"].dt == 0.25


def _round_trip(tmp, backend):
    system, p, kw, param = _prepare(tmp)
    with backend():
        assert ace_cli.main([param]) == 0
        direct = system(0, 8.0, p, **kw)
    data = np.genfromtxt(tmp + "b_linear_k7.out")          # the reference's own reader (:342, :104-110)
    parsed = gs.read_result(data, 3)
    assert parsed.shape == direct.shape == (4, 33)
    scale = np.maximum(np.abs(direct), 1e-30)
    assert (np.abs(parsed - direct) / scale).max() < 5e-11     # 12 significant digits per component
    return parsed


def test_cli_round_trip_on_oracle_backend(tmp_path):
    _round_trip(str(tmp_path) + "/", oracle_backend)


def test_cli_errors_exit_nonzero(tmp_path, capsys):
    bad = tmp_path / "bad.param"
    bad.write_text("dt 0.1\nta 0\nte 1\nadd_Output { |0><0|_2 + }\noutfile x.out\n")
    assert ace_cli.main([str(bad)]) == 1 and "ACE (aceqd-b200)" in capsys.readouterr().err
    assert ace_cli.main([]) == 2


def test_pt_generation_file_and_phonon_run(tmp_path):
    """Generation file as the reference writes it (:162-190), then a propagation that attaches the PT."""
    tmp = str(tmp_path) + "/"
    gen = tmp + "gen.param"
    with open(gen, "w") as fh:
        fh.write("dt 0.5\nte 4.0\nthreshold 1e-5\nuse_Gaussian_infinite true\ninfinite_normalize_iter 200\n"
                 "Boson_subtract_polaron_shift true\nBoson_E_min 0\nBoson_E_max 7\nBoson_SysOp {{ 1.000*|1><1|_2 }}\n"
                 "Boson_J_type QDPhonon\nBoson_J_a_e 5.0\ntemperature 4\ndont_propagate true\n"
                 "write_PT {}tls.pt\n".format(tmp))
    assert ace_cli.main([gen]) == 0 and os.path.exists(tmp + "tls.pt_initial")
    from pyaceqd_b200.two_level_system.tls import tls
    p = ChirpedPulse(tau_0=1.0, e_start=0, alpha=0, t0=3.0, e0=1.0)
    kw = dict(dt=0.5, lindblad=True, phonons=True, pt_file=tmp + "tls.pt", temp_dir=tmp, suffix="ph")
    tls(0, 6.0, p, prepare_only=True, **kw)
    with oracle_backend():
        assert ace_cli.main([tmp + "tls_ph.param"]) == 0
        direct = tls(0, 6.0, p, **kw)
    parsed = gs.read_result(np.genfromtxt(tmp + "tls_ph.out"), 4)
    assert np.abs(parsed - direct).max() < 1e-10
    assert abs(parsed[1, -1] + parsed[2, -1] - 1.0) < 1e-4        # trace preserved to the PT truncation level (1e-5)


@pytest.mark.gpu
def test_cli_subprocess_round_trip_on_gpu(tmp_path):
    """K7: the executable as a subprocess, exactly how the reference invokes ACE (:339-341)."""
    import contextlib
    tmp = str(tmp_path) + "/"
    system, p, kw, param = _prepare(tmp)
    subprocess.check_output([sys.executable, os.path.join(ROOT, "scripts", "ACE"), param])
    direct = system(0, 8.0, p, **kw)
    parsed = gs.read_result(np.genfromtxt(tmp + "b_linear_k7.out"), 3)
    assert (np.abs(parsed - direct) / np.maximum(np.abs(direct), 1e-30)).max() < 5e-11
    _round_trip(tmp, contextlib.nullcontext)


def test_aceutils_surface_runs_the_reference_dynmap_branch(tmp_path):
    """The object sequence of general_system.py:313-336 on the stand-in module."""
    from pyaceqd_b200 import ACEutils as au
    from pyaceqd_b200.two_level_system.tls import tls
    tmp = str(tmp_path) + "/"
    p = ChirpedPulse(tau_0=1.0, e_start=0.2, alpha=0, t0=3.0, e0=1.5)
    tls(0, 6.0, p, dt=0.25, lindblad=True, temp_dir=tmp, suffix="dm", prepare_only=True)
    with oracle_backend():
        plist = open(tmp + "tls_dm.param").readlines()
        param = au.Parameters(plist)
        initial_state = au.InitialState(param)
        fprop = au.FreePropagator(param)
        fprop.update(1.0, 0.25)
        PT, outp, tgrid, sim = au.ProcessTensors(param), au.OutputPrinter(param), au.TimeGrid(param), au.Simulation(param)
        sim.run(fprop, PT, initial_state, tgrid, outp)
        dm = np.array(au.DynamicalMap(fprop, PT, sim, tgrid).E)
        direct, E = tls(0, 6.0, p, dt=0.25, lindblad=True, calc_dynmap=True)
        M = tls(0, 6.0, p, dt=0.25, lindblad=True, get_M_t=1.0)
    data = np.genfromtxt(tmp + "tls_dm.out", usecols=list(range(9)))
    assert np.abs(gs.read_result(data, 4) - direct).max() < 1e-10
    assert dm.shape == (24, 4, 4) and np.abs(dm - E).max() < 1e-12
    assert np.abs(fprop.M - M).max() < 1e-12
    rho_t = dm @ np.array([1, 0, 0, 0], dtype=complex)
    assert np.abs(rho_t[:, 3] - direct[2][1:]).max() < 1e-10
    other = au.InitialState(np.array([[0, 0], [0, 1.0]]))
    assert np.allclose(other.rho, [0, 0, 0, 1])
