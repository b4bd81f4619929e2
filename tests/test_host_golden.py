"""Host helpers against golden vectors generated from the reference itself
(tests/golden/make_reference_fixtures.py imports /root/reference in the build container)."""
import json
import os

import numpy as np
import pytest

from pyaceqd_b200 import pulses, tools

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_host.json")))


@pytest.mark.parametrize("key", sorted(G["pulses"]))
def test_pulse_fields_match_reference(key):
    spec = G["pulses"][key]
    p = getattr(pulses, spec["cls"])(**spec["kw"])
    t = np.asarray(G["pulse_t"])
    ref = np.asarray(spec["re"]) + 1j * np.asarray(spec["im"])
    assert np.abs(p.get_total(t) - ref).max() < 1e-14
    assert abs(p.polar_y - spec["polar_y"]) < 1e-15
    if spec["freq"] is not None:
        assert np.abs(np.asarray(p.get_frequency(t)) - np.asarray(spec["freq"])).max() < 1e-13


def test_pulse_copy_and_energy_roundtrip():
    p = pulses.ChirpedPulse(tau_0=3, e_start=-2, alpha=20, t0=12, e0=5, polar_x=0.6)
    q = p.copy()
    assert isinstance(q, pulses.ChirpedPulse) and q.get_energy() == p.get_energy()
    q.set_energy(0.0, 0.0)
    assert p.get_energy() != q.get_energy()
    assert abs(p.polar_x ** 2 + p.polar_y ** 2 - 1) < 1e-15


def test_pulse_area_normalisation():
    """int envelope dt = e0 (reference pulses.py:38-39)."""
    t = np.linspace(-40, 40, 80001)
    p = pulses.Pulse(tau=3.0, e_start=0.0, t0=0.0, e0=1.7)
    assert abs(np.trapezoid(p.get_envelope(t), t) - 1.7) < 1e-9
    c = pulses.ChirpedPulse(tau_0=3.0, e_start=0.0, alpha=0.0, t0=0.0, e0=1.0)
    assert abs(np.trapezoid(np.abs(c.get_total(t)), t) - 1.0) < 1e-9


def test_merge_intervals_reference_cases():
    for c in G["merge"]:   # pyaceqd/tests/test_merge_interval.py:5-23
        assert tools._merge_intervals([list(x) for x in c["inp"]]) == c["out"]


def test_time_grids_match_reference():
    p1 = pulses.ChirpedPulse(tau_0=1, e_start=0, t0=4)
    p2 = pulses.ChirpedPulse(tau_0=1, e_start=0, t0=20)
    p3 = pulses.ChirpedPulse(tau_0=1, e_start=0, t0=5)
    ct = G["construct_t"]
    np.testing.assert_allclose(tools.construct_t(0, 80, 0.1, 1.0, None, p1, p2), ct["two_pulses"], atol=1e-12)
    # reference quirk: the first pulse binds to dt_exp
    np.testing.assert_allclose(tools.construct_t(0, 80, 0.1, 1.0, p1, p2), ct["quirk_first_pulse_is_dt_exp"], atol=1e-12)
    np.testing.assert_allclose(tools.construct_t(0, 80, 0.1, 1.0, 0.1, p1, p3, simple_exp=True), ct["simple_exp"], atol=1e-12)
    np.testing.assert_allclose(tools.construct_t(0, 80, 0.1, 1.0, 0.1, p1, simple_exp=True, gaussian_t=True),
                               ct["gaussian_t"], atol=1e-12)
    np.testing.assert_allclose(tools.simple_t_gaussian(0, 10, 80, 0.1, 1.0, p1), G["simple_t_gaussian"], atol=1e-12)
    np.testing.assert_allclose(tools.round_to_dt(np.array([0.04, 0.11, 0.12, 0.26, 0.31]), 0.1), G["round_to_dt"], atol=1e-15)


def test_output_ops_dm_strings():
    """pyaceqd/tests/test_output_ops.py:11-24,43-73"""
    assert tools.output_ops_dm(2) == ["|0><0|_2", "|0><1|_2", "|1><1|_2"]
    for k, v in G["output_ops_dm"].items():
        dim = [int(x) for x in k.split("_")]
        assert tools.output_ops_dm(dim if len(dim) > 1 else dim[0]) == v
    assert len(tools.output_ops_dm(6)) == 21


def test_compose_dm_layout():
    """pyaceqd/tests/test_output_ops.py:26-41 and a random 3-level case from the reference."""
    data = np.zeros((4, 1), dtype=complex)
    data[1, 0], data[2, 0], data[3, 0] = 1, 3 + 3j, 2
    t, rho = tools.compose_dm(data, 2)
    assert np.allclose(rho[0], [[1, 3 + 3j], [3 - 3j, 2]]) and np.allclose(t, [0])
    c = G["compose_dm"]
    t, rho = tools.compose_dm(np.asarray(c["re"]) + 1j * np.asarray(c["im"]), dim=3)
    assert np.allclose(rho, np.asarray(c["rho_re"]) + 1j * np.asarray(c["rho_im"]))
    assert np.allclose(t, c["t"])


def test_concurrence_and_op_to_matrix():
    c = G["concurrence"]
    assert abs(tools.concurrence(np.asarray(c["re"]) + 1j * np.asarray(c["im"])) - c["value"]) < 1e-12
    bell = np.zeros((4, 4), dtype=complex)
    bell[0, 0] = bell[3, 3] = bell[0, 3] = bell[3, 0] = 0.5
    assert abs(tools.concurrence(bell) - G["concurrence_bell"]) < 1e-12
    assert np.allclose(tools.op_to_matrix("(|1><0|_3)"), G["op_to_matrix"])
    with pytest.raises(ValueError):
        tools.op_to_matrix("|3><0|_3")


def test_export_csv_format(tmp_path):
    f = tmp_path / "p.dat"
    tools.export_csv(str(f), np.array([0.0, 0.1]), np.array([1.0, 0.123456789]), np.array([0.0, -2.0]),
                     precision=8, delimit=" ")
    assert f.read_text().splitlines() == ["0.00000000 1.00000000 0.00000000", "0.10000000 0.12345679 -2.00000000"]
