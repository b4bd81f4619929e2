"""Operator-string grammar: every token the reference adapters emit."""
import numpy as np
import pytest

from pyaceqd_b200 import constants
from pyaceqd_b200.opparser import OperatorSyntaxError, parse_operator
from pyaceqd_b200.problem import build_problem, coupling_classes, liouville_left, liouville_right


def ketbra(i, j, d):
    m = np.zeros((d, d), complex)
    m[i, j] = 1
    return m


def test_basic_tokens():
    assert np.allclose(parse_operator("|1><0|_2"), ketbra(1, 0, 2))
    assert np.allclose(parse_operator("Id_3"), np.eye(3))
    b = parse_operator("b_3")
    assert np.allclose(b, np.diag([1, np.sqrt(2)], 1))
    assert np.allclose(parse_operator("bdagger_3"), b.conj().T)
    assert np.allclose(parse_operator("n_4"), np.diag([0, 1, 2, 3]))
    assert np.allclose(parse_operator("bdagger_3*b_3"), np.diag([0, 1, 2]))


def test_scalars_signs_and_constants():
    assert np.allclose(parse_operator("-0.5*pi*hbar*(|1><0|_2)"), -0.5 * np.pi * constants.hbar * ketbra(1, 0, 2))
    assert np.allclose(parse_operator("--4*|3><3|_4"), 4 * ketbra(3, 3, 4))       # linear.py:60 double sign
    assert np.allclose(parse_operator("-i*0.3*(|2><1|_6 - |1><2|_6 )"), -0.3j * (ketbra(2, 1, 6) - ketbra(1, 2, 6)))
    assert np.allclose(parse_operator("sqrt(2)*|0><1|_2"), np.sqrt(2) * ketbra(0, 1, 2))
    assert np.allclose(parse_operator("1e-3*|0><0|_2 + 2.5E+1*|1><1|_2"), np.diag([1e-3, 25.0]))
    assert np.allclose(parse_operator("0*|1><1|_2"), np.zeros((2, 2)))
    assert np.allclose(parse_operator("({}*|1><1|_2)".format(-0.0)), np.zeros((2, 2)))
    assert np.allclose(parse_operator("|1><1|_2/4"), 0.25 * ketbra(1, 1, 2))


def test_products_and_kron():
    a, bb, c = "|1><0|_2", "|1><1|_2", "|0><1|_2"
    assert np.allclose(parse_operator("(" + a + "*" + bb + "*" + c + ")"), ketbra(1, 0, 2) @ ketbra(1, 1, 2) @ ketbra(0, 1, 2))
    assert np.allclose(parse_operator("|1><0|_2 * |0><1|_2"), ketbra(1, 1, 2))
    k = parse_operator("|1><1|_2 otimes Id_2 otimes b_3")
    assert k.shape == (12, 12)
    assert np.allclose(k, np.kron(np.kron(ketbra(1, 1, 2), np.eye(2)), np.diag([1, np.sqrt(2)], 1)))
    # scalar * kron is unambiguous whichever way it associates
    assert np.allclose(parse_operator("2*|3><3|_4 otimes Id_2"), 2 * np.kron(ketbra(3, 3, 4), np.eye(2)))
    lo = parse_operator("0.1*(Id_2 otimes n_3) + 0.2*(|1><1|_2 otimes bdagger_3 + |1><1|_2 otimes b_3)")
    assert lo.shape == (6, 6) and np.allclose(lo, lo.conj().T)


def test_errors():
    for bad in ("|2><0|_2", "|1><0|_2 + |1><0|_3", "|1><0|_2 +", "foo_2", "3 + |0><0|_2", "|0><0|_2 * |0><0|_3"):
        with pytest.raises(OperatorSyntaxError):
            parse_operator(bad)
    with pytest.raises(OperatorSyntaxError):
        parse_operator("1.5")            # scalar without a dimension
    assert np.allclose(parse_operator("1.5", dim=2), 1.5 * np.eye(2))


def test_coupling_classes_match_survey_counts():
    """TLS {0,1} -> 4 classes; biexciton {0,1,1,2} -> 9; six-level -> 9."""
    assert len(coupling_classes(np.array([0.0, 1.0]))[1]) == 4
    cls, keys = coupling_classes(np.array([0.0, 1.0, 1.0, 2.0]))
    assert len(keys) == 9 and cls.shape == (16,)
    assert np.bincount(cls).tolist() == [1, 2, 1, 2, 4, 2, 1, 2, 1]
    assert len(coupling_classes(np.array([0, 1, 1, 1, 1, 2.0]))[1]) == 9
    assert len(coupling_classes(np.array([0.0, 0.0]))[1]) == 1   # no coupling -> one class


def test_liouville_conventions():
    """row-major vec: (A rho) <-> kron(A, I), (rho A) <-> kron(I, A^T); Tr(O rho) = w . vec(rho)."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    rho = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    assert np.allclose(liouville_left(a) @ rho.reshape(-1), (a @ rho).reshape(-1))
    assert np.allclose(liouville_right(a) @ rho.reshape(-1), (rho @ a).reshape(-1))
    p = build_problem(initial="|0><0|_2", output_ops=["|0><1|_2"], interaction_ops=[["|1><0|_2", "x"]])
    r = np.array([[0.3, 0.1 + 0.2j], [0.1 - 0.2j, 0.7]])
    assert abs(p.out_w[0] @ r.reshape(-1) - np.trace(ketbra(0, 1, 2) @ r)) < 1e-15   # <|0><1|> = rho_10
    # generator is trace preserving and Hermiticity preserving
    p2 = build_problem(system_op=["0.7*|1><1|_2"], initial="|0><0|_2", lindblad_ops=[["|0><1|_2", 0.3]],
                       interaction_ops=[["|1><0|_2", "x"]], output_ops=["|1><1|_2"])
    L = p2.liouvillian([0.2 - 0.1j])
    tr = np.eye(2).reshape(-1)
    assert np.abs(tr @ L).max() < 1e-14
