"""Named single-qubit kets -- drop-in for ``simulators/dv_simulator/states.py``
(enum members :5-12, ``get`` :17-31)."""
from __future__ import annotations

from enum import Enum

import numpy as np

from . import numpy_quantum as npq

_EIGHTH = np.pi / 8.0

# name -> amplitudes (length 2).  ZERO/ONE are integer arrays, PLUS/MINUS/H real, T/TDG
# complex -- the reference's dtypes, which matter because results follow NumPy promotion.
_AMPLITUDES = {
    "ZERO": lambda: npq.ZERO,
    "ONE": lambda: npq.ONE,
    "PLUS": lambda: npq.PLUS,
    "MINUS": lambda: npq.MINUS,
    "T": lambda: np.array([1.0, np.exp(2.0j * _EIGHTH)]) * 2 ** -0.5,
    "TDG": lambda: np.array([1.0, np.exp(-2.0j * _EIGHTH)]) * 2 ** -0.5,
    "H": lambda: np.array([np.cos(_EIGHTH), np.sin(_EIGHTH)]),
}


class _NamedKet(Enum):
    def __repr__(self):
        return self.name

    def get(self) -> np.ndarray:
        return _AMPLITUDES[self.name]()


State = _NamedKet("State", list(_AMPLITUDES), module=__name__)
