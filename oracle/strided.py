"""Strided (O(2^N) per gate) restatement of the same path as ``dense_ref``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Mathematically identical to the reference's dense-operator path
(DV/gates.py:44-54, :145-153, :165-186) but contracts the small matrix
directly with the target axes of the state tensor, so it finishes in seconds
up to ~24 qubits.  ``tests/test_oracle_golden.py`` pins it against the golden
vectors produced by the real reference and against ``dense_ref``.
Same plain-tuple op format and conventions as ``dense_ref``.
"""
from __future__ import annotations

import numpy as np

from .dense_ref import draw_outcome, measurement_vectors, qubit_count


def _contract(tensor: np.ndarray, matrix: np.ndarray, axes) -> np.ndarray:
    """Apply a (2^k, 2^k) matrix to ``axes`` of a rank-m tensor of 2-dims.
    Matrix factor f (most significant first) acts on ``axes[f]``."""
    k = len(axes)
    m = np.asarray(matrix).reshape((2,) * (2 * k))
    out = np.tensordot(m, tensor, axes=(list(range(k, 2 * k)), list(axes)))
    # tensordot puts the k new axes first; move them back to ``axes``
    return np.moveaxis(out, list(range(k)), list(axes))


def apply_matrix(state: np.ndarray, qubits, matrix: np.ndarray) -> np.ndarray:
    n = qubit_count(state)
    qubits = list(qubits)
    matrix = np.asarray(matrix)
    if state.ndim == 1:
        t = state.reshape((2,) * n) if n else state
        if n == 0:
            raise ValueError("no qubits to act on")
        return np.ascontiguousarray(_contract(t, matrix, qubits)).reshape(-1)
    if state.ndim == 2:
        t = state.reshape((2,) * (2 * n))
        t = _contract(t, matrix, qubits)
        t = _contract(t, np.conjugate(matrix), [q + n for q in qubits])
        return np.ascontiguousarray(t).reshape(state.shape)
    raise ValueError("state must be a ket or a density matrix")


def measure(state: np.ndarray, qubit: int, theta: float, phi: float,
            forced=None, uniform=None):
    n = qubit_count(state)
    v0, v1 = measurement_vectors(theta, phi)
    t = state.reshape((2,) * n)
    # a 1-qubit register collapses to a 0-d array in the reference (vector @ vector)
    shape = () if n == 1 else (-1,)
    r0 = np.tensordot(v0, t, axes=([0], [qubit])).reshape(shape)
    r1 = np.tensordot(v1, t, axes=([0], [qubit])).reshape(shape)
    n0 = np.linalg.norm(r0)
    n1 = np.linalg.norm(r1)
    if forced is not None:
        s = forced
    elif uniform is None:
        s = int(np.random.choice([0, 1], p=[n0 ** 2, n1 ** 2]))
    else:
        s = draw_outcome(n0 ** 2, n1 ** 2, uniform)
    return [r0, r1][s] / [n0, n1][s], s


def insert_qubit(state: np.ndarray, position: int, vec2: np.ndarray) -> np.ndarray:
    n = qubit_count(state)
    t = state.reshape((2,) * n) if n else state.reshape(())
    grown = np.multiply.outer(t, np.asarray(vec2))        # new axis last
    grown = np.moveaxis(grown, -1, position)
    return np.ascontiguousarray(grown).reshape(-1)


def apply_kraus(rho: np.ndarray, qubits, kraus, weights=None) -> np.ndarray:
    n = qubit_count(rho)
    qubits = list(qubits)
    t = rho.reshape((2,) * (2 * n))
    total = np.zeros(t.shape, dtype=np.complex128)
    for i, k in enumerate(kraus):
        k = np.asarray(k)
        term = _contract(_contract(t, k, qubits), np.conjugate(k), [q + n for q in qubits])
        total = total + (term if weights is None else weights[i] * term)
    return np.ascontiguousarray(total).reshape(rho.shape)


def run(ops, state: np.ndarray, uniforms=None):
    outcomes = []
    it = iter(uniforms) if uniforms is not None else None
    for op in ops:
        tag = op[0]
        if tag == "u":
            state = apply_matrix(state, op[1], op[2])
        elif tag == "m":
            u = next(it) if (it is not None and op[4] is None) else None
            state, s = measure(state, op[1], op[2], op[3], op[4], u)
            outcomes.append(s)
        elif tag == "ins":
            state = insert_qubit(state, op[1], op[2])
        elif tag == "kraus":
            state = apply_kraus(state, op[1], op[2])
        else:
            raise ValueError(f"unknown op tag {tag!r}")
    return state, outcomes
