"""Trajectory descriptions shared by the engine front-ends.

One :class:`Job` is what one ``system(t_start, t_end, *pulses, ...)`` call of the reference
asks ACE to do (``pyaceqd/general_system/general_system.py:128-360``): a time window, the
sampled drive tables (the pulse files of ``:55-71`` / rf file of ``:73-102``) and the
multi-time operator list (``:281-286``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .problem import MTO


@dataclass
class FieldTable:
    """A drive sampled on ``t0 + j*dt`` -- the content of one ACE pulse file."""
    t0: float
    dt: float
    values: np.ndarray  # complex128 [n]

    def __post_init__(self):
        self.values = np.ascontiguousarray(self.values, dtype=complex)


@dataclass
class Job:
    t_start: float
    t_end: float
    dt: float
    tables: Dict[str, FieldTable] = field(default_factory=dict)  # "x", "y", "rf"
    mtos: List[MTO] = field(default_factory=list)
    tail_rows: int = 0   # > 0: only the last `tail_rows` output rows are needed (0 = all)
    rho0: Optional[np.ndarray] = None   # initial vectorised state overriding the problem's (dynamical maps)
    table_len: int = 0   # > 0: this run's own drive has only this many samples (end value held beyond): the reference
                         # samples the pulse file of every run on np.arange(t_start, t_end, dt); 0 = the whole table

    @property
    def n_steps(self) -> int:
        # ACE: N = round((te - ta)/dt) steps -> N+1 output rows
        return int(round((self.t_end - self.t_start) / self.dt))

    def times(self) -> np.ndarray:
        return self.t_start + self.dt * np.arange(self.n_steps + 1)

    def mto_step(self, m: MTO) -> int:
        k = int(round((m.time - self.t_start) / self.dt))
        if abs(self.t_start + k * self.dt - m.time) > 1e-6 * max(1.0, abs(self.dt)):
            raise ValueError(f"multitime operator time {m.time} is not on the dt grid")
        if k < 0 or k > self.n_steps:
            raise ValueError(f"multitime operator time {m.time} outside [{self.t_start}, {self.t_end}]")
        return k
