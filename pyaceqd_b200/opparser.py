"""Operator-string grammar -> dense matrices.

In-process replacement for the part of ACE's parameter parser that the reference
feeds with the strings written at ``pyaceqd/general_system/general_system.py:239-289``
(``initial``, ``add_Hamiltonian``, ``add_Pulse``, ``add_Lindblad``, ``apply_Operator``,
``add_Output``).  Tokens actually produced by the reference adapters:

    |i><j|_d   Id_d   b_d   bdagger_d   n_d   otimes   + - * /   ( )   numbers
    i (imaginary unit)   hbar   pi   sqrt(x)

``otimes`` is a left-associative binary operator on the precedence level of ``*``
(Kronecker product, right factor fastest -- the ``itertools.product`` order of
``pyaceqd/tools.py:203-211``).  Python ``format`` artefacts such as ``--4*|3><3|_4``
(``four_level_system/linear.py:60``) parse as nested unary signs.
"""
from __future__ import annotations

import math
import re
from typing import List, Tuple, Union

import numpy as np

from . import constants

Value = Union[complex, np.ndarray]


class OperatorSyntaxError(ValueError):
    pass


_TOKEN = re.compile(
    r"""\s*(?:
      (?P<ketbra>\|\s*(\d+)\s*>\s*<\s*(\d+)\s*\|\s*_\s*(\d+)) |
      (?P<named>(?:Id|bdagger|b|n)_(\d+)) |
      (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?) |
      (?P<word>otimes|sqrt|hbar|pi|i\b) |
      (?P<sym>[-+*/()])
    )""",
    re.X,
)


def _tokenize(text: str) -> List[Tuple[str, object]]:
    toks: List[Tuple[str, object]] = []
    pos = 0
    n = len(text)
    while pos < n:
        if text[pos:].strip() == "":
            break
        m = _TOKEN.match(text, pos)
        if not m:
            raise OperatorSyntaxError(f"cannot parse operator string at {text[pos:pos+20]!r} in {text!r}")
        pos = m.end()
        if m.group("ketbra"):
            toks.append(("ketbra", (int(m.group(2)), int(m.group(3)), int(m.group(4)))))
        elif m.group("named"):
            name = m.group("named")
            kind, dim = name.rsplit("_", 1)
            toks.append(("named", (kind, int(dim))))
        elif m.group("num"):
            toks.append(("num", float(m.group("num"))))
        elif m.group("word"):
            toks.append(("word", m.group("word")))
        else:
            toks.append(("sym", m.group("sym")))
    toks.append(("end", None))
    return toks


def _ketbra(i: int, j: int, d: int) -> np.ndarray:
    if i >= d or j >= d:
        raise OperatorSyntaxError(f"|{i}><{j}|_{d}: index out of range")
    m = np.zeros((d, d), dtype=complex)
    m[i, j] = 1.0
    return m


def _named(kind: str, d: int) -> np.ndarray:
    if kind == "Id":
        return np.eye(d, dtype=complex)
    lower = np.diag(np.sqrt(np.arange(1, d, dtype=float)), 1).astype(complex)  # annihilator b
    if kind == "b":
        return lower
    if kind == "bdagger":
        return lower.conj().T
    if kind == "n":
        return np.diag(np.arange(d, dtype=float)).astype(complex)
    raise OperatorSyntaxError(kind)


class _Parser:
    def __init__(self, text: str):
        self.text = text
        self.toks = _tokenize(text)
        self.k = 0

    def peek(self):
        return self.toks[self.k]

    def take(self):
        t = self.toks[self.k]
        self.k += 1
        return t

    def expect_sym(self, s: str):
        t = self.take()
        if t != ("sym", s):
            raise OperatorSyntaxError(f"expected {s!r} in {self.text!r}")

    # expr := term (('+'|'-') term)*
    def expr(self) -> Value:
        v = self.term()
        while self.peek() in (("sym", "+"), ("sym", "-")):
            op = self.take()[1]
            r = self.term()
            v = _add(v, r if op == "+" else _neg(r), self.text)
        return v

    # term := unary (('*'|'/'|'otimes') unary)*
    def term(self) -> Value:
        v = self.unary()
        while True:
            t = self.peek()
            if t == ("sym", "*"):
                self.take()
                v = _mul(v, self.unary(), self.text)
            elif t == ("sym", "/"):
                self.take()
                r = self.unary()
                if isinstance(r, np.ndarray):
                    raise OperatorSyntaxError(f"division by an operator in {self.text!r}")
                v = v / r
            elif t == ("word", "otimes"):
                self.take()
                v = _kron(v, self.unary())
            else:
                return v

    def unary(self) -> Value:
        t = self.peek()
        if t == ("sym", "-"):
            self.take()
            return _neg(self.unary())
        if t == ("sym", "+"):
            self.take()
            return self.unary()
        return self.atom()

    def atom(self) -> Value:
        kind, val = self.take()
        if kind == "num":
            return complex(val)
        if kind == "ketbra":
            return _ketbra(*val)
        if kind == "named":
            return _named(*val)
        if kind == "word":
            if val == "i":
                return 1j
            if val == "hbar":
                return complex(constants.hbar)
            if val == "pi":
                return complex(math.pi)
            if val == "sqrt":
                self.expect_sym("(")
                v = self.expr()
                self.expect_sym(")")
                if isinstance(v, np.ndarray):
                    raise OperatorSyntaxError("sqrt of an operator")
                return complex(np.sqrt(v))
        if (kind, val) == ("sym", "("):
            v = self.expr()
            self.expect_sym(")")
            return v
        raise OperatorSyntaxError(f"unexpected token {val!r} in {self.text!r}")


def _neg(v: Value) -> Value:
    return -v


def _add(a: Value, b: Value, text: str) -> Value:
    am, bm = isinstance(a, np.ndarray), isinstance(b, np.ndarray)
    if am and bm:
        if a.shape != b.shape:
            raise OperatorSyntaxError(f"dimension mismatch in sum: {a.shape} vs {b.shape} in {text!r}")
        return a + b
    if not am and not bm:
        return a + b
    # scalar + operator: only a literal zero is meaningful (e.g. "0*|0><0|_2 + 0")
    s, m = (a, b) if bm else (b, a)
    if s == 0:
        return m
    raise OperatorSyntaxError(f"scalar added to operator in {text!r}")


def _mul(a: Value, b: Value, text: str) -> Value:
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        if a.shape != b.shape:
            raise OperatorSyntaxError(f"dimension mismatch in product: {a.shape} vs {b.shape} in {text!r}")
        return a @ b
    return a * b


def _kron(a: Value, b: Value) -> Value:
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        return np.kron(a, b)
    return a * b


def parse_operator(text: str, dim: int | None = None) -> np.ndarray:
    """Parse one operator expression (the text between ``{ }`` of an ACE param line).

    Returns a dense complex ``[N, N]`` matrix.  ``dim`` (if given) is checked.
    A purely scalar expression is only accepted together with ``dim`` and is
    returned as ``scalar * Id`` (needed for e.g. ``0*...`` collapses).
    """
    p = _Parser(text)
    v = p.expr()
    if p.peek()[0] != "end":
        raise OperatorSyntaxError(f"trailing tokens in {text!r}")
    if not isinstance(v, np.ndarray):
        if dim is None:
            raise OperatorSyntaxError(f"{text!r} is a scalar; operator dimension unknown")
        v = complex(v) * np.eye(dim, dtype=complex)
    if dim is not None and v.shape != (dim, dim):
        raise OperatorSyntaxError(f"{text!r} has dimension {v.shape[0]}, expected {dim}")
    return np.ascontiguousarray(v, dtype=complex)


def operator_dim(text: str) -> int:
    """Hilbert-space dimension of an operator expression."""
    return parse_operator(text).shape[0]
