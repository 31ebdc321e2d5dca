"""Dense-operator restatement of the reference gate-application algorithm.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every routine follows, step by step, what the reference does in
``/root/reference/simulators/dv_simulator`` (abbreviated ``DV/`` below):
a k-qubit matrix is blown up to the full 2^N x 2^N operator by a chain of
Kronecker products, its tensor factors are moved into place by a
reshape/transpose/reshape, and the operator is multiplied onto the state.
Cost is O(4^N) memory and time per gate, exactly like the reference, which is
why this module is also the timed CPU baseline of ``bench.py``.

Conventions reproduced: qubit 0 is the most significant
index bit; the first tensor factor of a k-qubit matrix acts on ``qubits[0]``;
measurement contracts with the *un-conjugated* vector ``Rz(phi) Ry(theta) e_s``
and removes the qubit; insertion grows the register.

Circuits are given as plain tuples so the oracle does not depend on the product
package:

    ("u",     qubits, matrix)                     unitary / any square matrix
    ("m",     qubit, theta, phi, forced_or_None)  projective measurement
    ("ins",   position, vec2)                     qubit insertion
    ("kraus", qubits, [K_0, K_1, ...])            channel on a density matrix
"""
from __future__ import annotations

import math
from functools import reduce

import numpy as np

_ID2 = np.identity(2)
_E0 = np.array([1, 0])
_E1 = np.array([0, 1])
_SX = np.array([[0, 1], [1, 0]])
_SY = np.array([[0, -1j], [1j, 0]])
_SZ = np.array([[1, 0], [0, -1]])


def qubit_count(arr: np.ndarray) -> int:
    """log2 of the leading dimension, truncated like DV/numpy_quantum.py:254-258."""
    return int(np.log2(arr.shape[0]))


def kron_chain(*factors) -> np.ndarray:
    """Left-to-right Kronecker product seeded with the scalar 1
    (DV/numpy_quantum.py:169-170)."""
    return reduce(np.kron, factors, 1)


def _invert(order):
    inv = [0] * len(order)
    for src, dst in enumerate(order):
        inv[dst] = src
    return inv


def _move_row_factors(arr: np.ndarray, axes_order) -> np.ndarray:
    """Split the leading dimension into one axis per qubit, transpose, merge
    (DV/numpy_quantum.py:212-217)."""
    nq = len(axes_order)
    cube = arr.reshape((2,) * nq + (-1,))
    cube = cube.transpose(list(axes_order) + [nq])
    return cube.reshape((2 ** nq, -1))


def move_factors(arr: np.ndarray, destination) -> np.ndarray:
    """Tensor factor j of ``arr`` ends up at position ``destination[j]``; for
    operators both the row and the column space are permuted
    (DV/numpy_quantum.py:227-240)."""
    dim = arr.shape[0]
    if dim == 0 or dim & (dim - 1):
        raise ValueError("leading dimension is not a power of two")
    if set(destination) != set(range(qubit_count(arr))):
        raise ValueError("destination is not a permutation of all qubits")
    order = _invert(destination)
    out = _move_row_factors(arr, order)
    if arr.ndim == 2:
        out = _move_row_factors(out.T, order).T
    else:
        out = out.flatten()
    return out


def full_operator(matrix: np.ndarray, n: int, qubits) -> np.ndarray:
    """(2^n, 2^n) operator acting with ``matrix`` on ``qubits`` and identity
    elsewhere (DV/numpy_quantum.py:243-247)."""
    qubits = list(qubits)
    rest = [q for q in range(n) if q not in qubits]
    op = kron_chain(matrix, *([_ID2] * len(rest)))
    return move_factors(op, qubits + rest)


def apply_matrix(state: np.ndarray, qubits, matrix: np.ndarray) -> np.ndarray:
    """Ket: U psi.  Density matrix: U rho U^dagger (DV/gates.py:44-54)."""
    n = qubit_count(state)
    op = full_operator(matrix, n, qubits)
    if state.ndim == 1:
        return op @ state
    if state.ndim == 2:
        return op @ state @ np.conjugate(op.T)
    raise ValueError("state must be a ket or a density matrix")


def rotation(theta: float, axis) -> np.ndarray:
    """exp(-i theta/2 axis.sigma) (DV/numpy_quantum.py:104-105)."""
    gen = axis[0] * _SX + axis[1] * _SY + axis[2] * _SZ
    return _ID2 * np.cos(theta / 2) - 1j * gen * np.sin(theta / 2)


def measurement_vectors(theta: float, phi: float):
    """The two (un-conjugated) contraction vectors of DV/gates.py:169-171."""
    rot = rotation(phi, [0, 0, 1]) @ rotation(theta, [0, 1, 0])
    return rot @ _E0, rot @ _E1


def measure(state: np.ndarray, qubit: int, theta: float, phi: float,
            forced=None, uniform=None):
    """Projective measurement that removes the qubit (DV/gates.py:165-186).

    Builds the two dense (2^(N-1), 2^N) contraction operators, takes the
    norms, draws the outcome and returns ``(collapsed / norm, outcome)``.
    ``uniform`` supplies the random number explicitly; when it is ``None`` the
    draw comes from the global legacy NumPy generator exactly like the
    reference (``np.random.choice([0, 1], p=...)``).
    """
    n = qubit_count(state)
    v0, v1 = measurement_vectors(theta, phi)
    factors = [_ID2] * n
    factors[qubit] = v0
    r0 = kron_chain(*factors) @ state
    n0 = np.linalg.norm(r0)
    factors[qubit] = v1
    r1 = kron_chain(*factors) @ state
    n1 = np.linalg.norm(r1)
    if forced is not None:
        s = forced
    elif uniform is None:
        s = int(np.random.choice([0, 1], p=[n0 ** 2, n1 ** 2]))
    else:
        s = draw_outcome(n0 ** 2, n1 ** 2, uniform)
    return [r0, r1][s] / [n0, n1][s], s


def draw_outcome(p0: float, p1: float, uniform: float) -> int:
    """What ``np.random.choice([0,1], p=[p0,p1])`` returns for the uniform
    sample ``uniform`` (legacy RandomState: cdf = cumsum(p); cdf /= cdf[-1];
    searchsorted(cdf, u, side='right')), including its tolerance check."""
    if abs(p0 + p1 - 1.0) > math.sqrt(np.finfo(np.float64).eps):
        raise ValueError("probabilities do not sum to 1")
    cdf = np.cumsum(np.array([p0, p1], dtype=np.float64))
    cdf /= cdf[-1]
    return int(cdf.searchsorted(uniform, side="right"))


def insert_qubit(state: np.ndarray, position: int, vec2: np.ndarray) -> np.ndarray:
    """psi (x) chi, then move the new last factor to ``position``
    (DV/gates.py:145-153)."""
    n = qubit_count(state)
    grown = kron_chain(state, vec2)
    dest = list(range(position)) + list(range(position + 1, n + 1)) + [position]
    return move_factors(grown, dest)


def apply_kraus(rho: np.ndarray, qubits, kraus, weights=None) -> np.ndarray:
    """sum_i [w_i] K_i rho K_i^dagger with every K_i expanded to full size; the
    semantics of quantum_channel in
    impact_of_finite_squeezing_.../tomography.py:21-24."""
    n = qubit_count(rho)
    total = np.zeros(rho.shape, dtype=np.complex128)
    for i, k in enumerate(kraus):
        big = full_operator(np.asarray(k), n, qubits)
        term = big @ rho @ np.conjugate(big.T)
        total = total + (term if weights is None else weights[i] * term)
    return total


def product_state(vectors) -> np.ndarray:
    """Kronecker product of single-qubit kets (DV/simulator.py:26)."""
    return kron_chain(*vectors)


def run(ops, state: np.ndarray, uniforms=None):
    """Sequential executor (DV/simulator.py:36-53) over plain-tuple ops.
    Returns ``(final_state, outcomes)``.  ``uniforms`` is an iterator of
    explicit random numbers for measurements (``None`` = global NumPy RNG)."""
    outcomes = []
    it = iter(uniforms) if uniforms is not None else None
    for op in ops:
        tag = op[0]
        if tag == "u":
            state = apply_matrix(state, op[1], np.asarray(op[2]))
        elif tag == "m":
            u = next(it) if (it is not None and op[4] is None) else None
            state, s = measure(state, op[1], op[2], op[3], op[4], u)
            outcomes.append(s)
        elif tag == "ins":
            state = insert_qubit(state, op[1], np.asarray(op[2]))
        elif tag == "kraus":
            state = apply_kraus(state, op[1], op[2])
        else:
            raise ValueError(f"unknown op tag {tag!r}")
    return state, outcomes


# ---- state metrics (DV/numpy_quantum.py:131-166) ---------------------------

def fidelity(a: np.ndarray, b: np.ndarray) -> float:
    ka, kb = a.ndim == 1, b.ndim == 1
    if ka and kb:
        return np.abs(a.conj() @ b).real ** 2
    if ka:
        return (a.conj() @ b @ a).real
    if kb:
        return (b.conj() @ a @ b).real
    ev = np.clip(np.linalg.eigvals(a @ b).real, 0.0, None)
    return np.sum(np.sqrt(ev)) ** 2


def purity(rho: np.ndarray) -> float:
    return np.trace(rho @ rho).real
