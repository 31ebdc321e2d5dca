"""Deterministic synthetic circuits for the benchmark configurations
(BASELINE.json ``configs``).

Every generator takes the gate module to instantiate from (``gates=`` defaults
to this package's), so ``tests/golden/make_golden.py`` can build the *same*
circuit out of the reference's own classes and run it through the reference.
"""
from __future__ import annotations

import numpy as np


def _gates(gates):
    if gates is None:
        from . import gates as gates
    return gates


def sv_random_circuit(n: int, depth: int, seed: int, gates=None) -> list:
    """Config C4/C5: ``depth`` layers; each layer gives every qubit one gate drawn
    uniformly from {H, T, RZ(theta ~ U[0, 2pi)), X, P}, then CZ on a uniformly
    random perfect matching of the qubits (arbitrary distance; with odd ``n``
    one qubit sits out).  ``rng = np.random.default_rng(seed)``."""
    g = _gates(gates)
    rng = np.random.default_rng(seed)
    circuit = []
    for _ in range(depth):
        kinds = rng.integers(0, 5, size=n)
        angles = rng.uniform(0.0, 2.0 * np.pi, size=n)
        for q in range(n):
            k = int(kinds[q])
            if k == 0:
                circuit.append(g.H(q))
            elif k == 1:
                circuit.append(g.T(q))
            elif k == 2:
                circuit.append(g.RZ(q, float(angles[q])))
            elif k == 3:
                circuit.append(g.X(q))
            else:
                circuit.append(g.P(q))
        order = rng.permutation(n)
        for i in range(0, n - 1, 2):
            circuit.append(g.CZ(int(order[i]), int(order[i + 1])))
    return circuit


def inverse_circuit(circuit, gates=None) -> list:
    """Gate-by-gate inverse (reversed order, conjugate-transposed matrices) as
    generic ``Gate`` objects; used for the circuit-then-inverse check at sizes no
    oracle can reach."""
    g = _gates(gates)
    out = []
    for gate in reversed(circuit):
        out.append(g.Gate(list(gate.indices), np.conjugate(np.asarray(gate.matrix).T)))
    return out


def dm_random_layers(n: int, depth: int, seed: int, gates=None) -> list:
    """Config C3 without the noise: ``depth`` layers; every qubit gets one gate
    from {H, P, Pdg, T, Tdg}, then CZ on a random disjoint nearest-neighbour
    pairing (start offset 0 or 1, each candidate pair kept with probability 1/2).
    Returns a list of layers, each a list of gates, so the caller can interleave
    noise channels."""
    g = _gates(gates)
    rng = np.random.default_rng(seed)
    singles = (g.H, g.P, g.Pdg, g.T, g.Tdg)
    layers = []
    for _ in range(depth):
        layer = []
        kinds = rng.integers(0, len(singles), size=n)
        for q in range(n):
            layer.append(singles[int(kinds[q])](q))
        offset = int(rng.integers(0, 2))
        keep = rng.integers(0, 2, size=n)
        for a in range(offset, n - 1, 2):
            if keep[a]:
                layer.append(g.CZ(a, a + 1))
        layers.append(layer)
    return layers


# ---- randomised benchmarking (PAPER/randomised_benchmarking.py:27-49) ---------------------
RB_GATE_NAMES = ("I", "H", "P", "Pdg", "CZ", "SWAP")


def rb_random_circuit(N: int, depth: int, rng, gates=None) -> list:
    """The reference's ``random_circ`` (randomised_benchmarking.py:29-49) with the
    same draws from ``rng``: gates from (I, H, P, Pdg, CZ, SWAP) are appended until
    the measurement-based layering of the circuit reaches ``depth`` layers.

    The layer count is what ``MBGKPCircuit.depth()`` reports
    (GKP/transpiler.py:144, :188-198) for this gate set: every gate is placed in
    the layer right after the last one occupied on any of its qubits, and an
    empty circuit already counts one layer."""
    if N < 2:
        raise ValueError("At least 2 qubits required!")
    g = _gates(gates)
    classes = [getattr(g, name) for name in RB_GATE_NAMES]
    used = [0] * N                    # layers occupied so far on each qubit
    circuit = []
    while max(1, max(used)) < depth:
        cls = classes[int(rng.choice(len(classes), 1)[0])]
        if issubclass(cls, g.SingleQubitGate):
            i = int(rng.choice(N, 1)[0])
            circuit.append(cls(i))
            used[i] += 1
        else:
            i = int(rng.choice(N - 1, 1)[0])
            circuit.append(cls(i, i + 1))
            used[i] = used[i + 1] = max(used[i], used[i + 1]) + 1
    return circuit
