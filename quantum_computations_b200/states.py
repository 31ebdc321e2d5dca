"""Named single-qubit kets -- drop-in for ``simulators/dv_simulator/states.py``
(enum members :5-12, ``get`` :17-31)."""
from __future__ import annotations

from enum import Enum, auto

import numpy as np

from . import numpy_quantum as npq


class State(Enum):
    ZERO = auto()
    ONE = auto()
    PLUS = auto()
    MINUS = auto()
    T = auto()
    TDG = auto()
    H = auto()

    def __repr__(self):
        return self.name

    def get(self) -> np.ndarray:
        """Amplitudes (length 2).  ZERO/ONE are integer arrays, PLUS/MINUS/H
        real, T/TDG complex -- the reference's dtypes."""
        if self is State.ZERO:
            return npq.ZERO
        if self is State.ONE:
            return npq.ONE
        if self is State.PLUS:
            return npq.PLUS
        if self is State.MINUS:
            return npq.MINUS
        if self is State.T:
            return np.array([1.0, np.exp(1.0j * np.pi / 4.0)]) * 2 ** -0.5
        if self is State.TDG:
            return np.array([1.0, np.exp(-1.0j * np.pi / 4.0)]) * 2 ** -0.5
        return np.array([np.cos(np.pi / 8.0), np.sin(np.pi / 8.0)])
