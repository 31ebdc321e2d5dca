"""Analytic finite-squeezing GKP logical-error model (oracle side).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, independently of the product package, the closed-form model the
reference uses for its "Analytical estimate" curves:

* ``db2eps``:  eps = 2 atanh(10^(-dB/10) / 2)
  (simulators/gkp_simulator/utils.py:14-15)
* per-quadrature logical flip probability for an input quadrature variance of
  ``k * eps / 2``:  e(k) = 1 - erf( sqrt( pi / (8 k eps / 2) ) )
  (impact_of_finite_squeezing_.../plot_data.ipynb:64-68)
* I/H-type gate: both quadratures at k=2; P-type gate: k=2 and k=3
  (plot_data.ipynb:71-75).

The Kraus set of the resulting single-qubit Pauli channel (independent X flip
with probability ``px`` and Z flip with probability ``pz``) is what the parity
tests feed to ``dense_ref.apply_kraus``.
"""
from __future__ import annotations

import math

import numpy as np

_I = np.identity(2)
_X = np.array([[0.0, 1.0], [1.0, 0.0]])
_Z = np.array([[1.0, 0.0], [0.0, -1.0]])


def db2eps(db: float) -> float:
    return 2.0 * math.atanh(math.pow(10.0, -db / 10.0) / 2.0)


def flip_probability(db: float, k: int) -> float:
    var = k * db2eps(db) / 2.0
    return 1.0 - math.erf(math.sqrt(math.pi / (8.0 * var)))


def gate_error_I(db: float) -> float:
    e2 = flip_probability(db, 2)
    return 1.0 - (1.0 - e2) * (1.0 - e2)


def gate_error_P(db: float) -> float:
    return 1.0 - (1.0 - flip_probability(db, 2)) * (1.0 - flip_probability(db, 3))


def pauli_flip_kraus(px: float, pz: float):
    """Kraus operators of: X applied with probability px, then Z with pz."""
    return [
        math.sqrt((1 - px) * (1 - pz)) * _I,
        math.sqrt(px * (1 - pz)) * _X,
        math.sqrt((1 - px) * pz) * _Z,
        math.sqrt(px * pz) * (_Z @ _X),
    ]
