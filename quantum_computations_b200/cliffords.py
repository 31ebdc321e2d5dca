"""Two-qubit Clifford group modulo Paulis.

The reference enumerates the 720 classes of two-qubit Cliffords modulo the Pauli
group by a breadth-first search over their symplectic representations
(``PAPER/average_clifford_fidelity.py:65-151``: generators H, P on either qubit,
CX both ways, SWAP; 720 classes; Cayley-graph diameter 7).  This module does the
same on the host and turns the table into *uniform* Clifford RB sequences for the
batched GPU executor (``batched.BatchedSimulator``), upgrading configuration C2
from generator-set RB to true Clifford RB.
"""
from __future__ import annotations

from collections import deque
from functools import lru_cache

import numpy as np

from . import numpy_quantum as npq
from .gates import Gate

_I2 = np.identity(2, dtype=complex)
_P1 = {(0, 0): _I2, (1, 0): npq.X.astype(complex), (0, 1): npq.Z.astype(complex),
       (1, 1): (npq.Z @ npq.X).astype(complex)}


def pauli_from_bits(x1: int, z1: int, x2: int, z2: int) -> np.ndarray:
    """The two-qubit Pauli Z^z X^x on each qubit (phase-free representative)."""
    return np.kron(_P1[(x1, z1)], _P1[(x2, z2)])


def pauli_bits(op: np.ndarray):
    """(x1, z1, x2, z2) of a 4x4 matrix proportional to a Pauli operator."""
    for bits in np.ndindex(2, 2, 2, 2):
        if abs(np.trace(pauli_from_bits(*bits).conj().T @ op)) > 1e-9:
            return tuple(int(b) for b in bits)
    raise ValueError("operator is not proportional to a Pauli")


def symplectic_rep(unitary: np.ndarray) -> np.ndarray:
    """4x4 binary matrix whose columns are the images of X1, Z1, X2, Z2 under
    conjugation by ``unitary`` (phases dropped)."""
    cols = []
    for gen in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)):
        image = unitary @ pauli_from_bits(*gen) @ unitary.conj().T
        cols.append(pauli_bits(image))
    return np.array(cols, dtype=int).T % 2


def generators():
    h, p = npq.H.astype(complex), npq.P.astype(complex)
    cx = npq.CX.astype(complex)
    cx_rev = npq.permute_tensor_product(npq.CX, [1, 0]).astype(complex)
    return [np.kron(h, _I2), np.kron(_I2, h), np.kron(p, _I2), np.kron(_I2, p), cx, cx_rev, npq.SWAP.astype(complex)]


@lru_cache(maxsize=1)
def two_qubit_cliffords_mod_paulis():
    """(unitaries, depths): one unitary per symplectic class, found by BFS from the
    identity, and the word length at which each class is first reached."""
    gens = [(symplectic_rep(g), g) for g in generators()]
    start = np.eye(4, dtype=int)
    table = {start.tobytes(): (np.eye(4, dtype=complex), 0)}
    queue = deque([start])
    while queue:
        s = queue.popleft()
        u, d = table[s.tobytes()]
        for sg, ug in gens:
            s_new = (sg @ s) % 2
            key = s_new.tobytes()
            if key not in table:
                table[key] = (ug @ u, d + 1)
                queue.append(s_new)
    unitaries = [u for u, _ in table.values()]
    depths = [d for _, d in table.values()]
    return unitaries, depths


def random_clifford(rng: np.random.Generator) -> np.ndarray:
    """A uniformly random two-qubit Clifford (class representative times a uniformly
    random Pauli), as a 4x4 unitary."""
    unitaries, _ = two_qubit_cliffords_mod_paulis()
    u = unitaries[int(rng.integers(0, len(unitaries)))]
    bits = rng.integers(0, 2, size=4)
    return pauli_from_bits(*bits) @ u


def clifford_rb_sequences(num_sequences: int, length: int, rng: np.random.Generator, *, invert: bool = True):
    """``num_sequences`` circuits of ``length`` uniformly random Cliffords on qubits
    (0, 1), each followed (``invert``) by the Clifford that undoes the sequence.

    A Clifford is emitted as its class representative (one of 720 ``Gate([0, 1], U)``
    objects, shared between sequences so the batched executor needs only 720 + 4
    opcodes) followed by the X / Z gates of a uniformly random Pauli; the final
    inverse is one generic gate."""
    from . import gates as _gates
    unitaries, _ = two_qubit_cliffords_mod_paulis()
    class_gates = [Gate([0, 1], u) for u in unitaries]
    paulis = {(0, "x"): _gates.X(0), (0, "z"): _gates.Z(0), (1, "x"): _gates.X(1), (1, "z"): _gates.Z(1)}
    out = []
    for _ in range(num_sequences):
        total = np.eye(4, dtype=complex)
        circ = []
        for _ in range(length):
            k = int(rng.integers(0, len(class_gates)))
            x1, z1, x2, z2 = (int(b) for b in rng.integers(0, 2, size=4))
            circ.append(class_gates[k])
            total = unitaries[k] @ total
            for q, kind, on in ((0, "x", x1), (0, "z", z1), (1, "x", x2), (1, "z", z2)):
                if on:
                    circ.append(paulis[(q, kind)])
            total = pauli_from_bits(x1, z1, x2, z2) @ total
        if invert:
            circ.append(Gate([0, 1], total.conj().T))
        out.append(circ)
    return out
