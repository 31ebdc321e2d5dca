"""Measurement-based layer scheduling of a DV circuit.

The reference's GKP simulator places every DV gate into *layers* of a
measurement-based computation (``GKP/transpiler.py:65-209``): a gate goes into the
layer right after the last layer in which any of its qubits is busy, Pauli gates
are not executed but recorded as byproducts on the layer they follow, a ``T`` /
``Tdg`` drags a classically controlled ``P`` / ``Pdg`` correction behind it, and
``fill()`` pads idle qubits with identity gates (which matter: every teleportation
step adds noise, also on idle qubits).  ``random_circ`` of the RB script stops
adding gates once ``depth()`` reaches its target
(``PAPER/randomised_benchmarking.py:29-49``).

Here the same bookkeeping is a small host-side scheduler whose layers are natural
fusion units for the GPU planner and the unit of noise accounting
(``noisy_circuit``).  Only the DV side is rebuilt; the CV gate objects of the
reference (``gate_transpile``) are out of scope.
"""
from __future__ import annotations

from . import gates as _g
from .simulator import ClassicalControl

IMPLEMENTABLE = ("I", "H", "P", "Pdg", "T", "Tdg", "CZ", "SWAP")
PAULI_BITS = {"X": (1, 0), "Y": (1, 1), "Z": (0, 1)}     # (x, z) byproduct of each Pauli gate


class Layer:
    def __init__(self, num_qubits: int):
        self.num_qubits = num_qubits
        self.busy = [False] * num_qubits
        self.gates = []                                    # kept sorted by lowest qubit index
        self.paulis = [[0, 0] for _ in range(num_qubits)]  # (x, z) byproducts recorded on this layer

    def gate_on(self, qubit: int):
        for gate in self.gates:
            if qubit in gate.indices:
                return gate
        return None

    def occupied(self, qubits) -> bool:
        return any(self.busy[q] or self.paulis[q] != [0, 0] for q in qubits)

    def place(self, gate) -> None:
        for q in gate.indices:
            self.busy[q] = True
        key = min(gate.indices)
        pos = len(self.gates)
        while pos > 0 and min(self.gates[pos - 1].indices) > key:
            pos -= 1
        self.gates.insert(pos, gate)

    def fill(self) -> None:
        for q in range(self.num_qubits):
            if self.gate_on(q) is None:
                self.place(_g.I(q))


class MBLayering:
    """Layer structure of a circuit on ``num_qubits`` qubits (nearest-neighbour
    two-qubit gates only, like the reference)."""

    def __init__(self, num_qubits: int):
        self.num_qubits = num_qubits
        self.layers = [Layer(num_qubits)]

    @classmethod
    def of(cls, circuit, num_qubits: int | None = None) -> "MBLayering":
        if num_qubits is None:
            num_qubits = max(max(g.indices) for g in circuit) + 1
        out = cls(num_qubits)
        for gate in circuit:
            out.add_gate(gate)
        return out

    def depth(self) -> int:
        return len(self.layers)

    def count(self) -> int:
        return sum(len(layer.gates) for layer in self.layers)

    def fill(self) -> None:
        for layer in self.layers:
            layer.fill()

    # -- construction -------------------------------------------------------------------------
    def _last_occupied(self, qubits):
        """Index (from the back, -1 = last) of the latest layer busy on any of the qubits."""
        for back in range(1, len(self.layers) + 1):
            if self.layers[-back].occupied(qubits):
                return -back
        return None

    def _schedule(self, gate) -> None:
        at = self._last_occupied(gate.indices)
        if at is None:
            target = self.layers[0]
        elif at == -1:
            self.layers.append(Layer(self.num_qubits))
            target = self.layers[-1]
        else:
            target = self.layers[at + 1]
        target.place(gate)

    def add_gate(self, gate) -> None:
        idx = gate.indices
        if any(q < 0 or q >= self.num_qubits for q in idx):
            raise ValueError(f"Cannot add {gate} to MBGKPCircuit with {self.num_qubits} qubits.")
        if len(idx) > 2:
            raise ValueError(f"Only single- and two-mode gates available, but gate {gate} was given.")
        if len(idx) == 2 and abs(idx[0] - idx[1]) != 1:
            raise ValueError(f"Only nearest neighbour interactions available, but gate {gate} was given.")
        name = type(gate).__name__
        if name in IMPLEMENTABLE:
            self._schedule(gate)
            if name == "T":
                self._schedule(ClassicalControl(_g.P(idx[0]), [-self.num_qubits]))
            elif name == "Tdg":
                self._schedule(ClassicalControl(_g.Pdg(idx[0]), [-self.num_qubits]))
        elif name in PAULI_BITS:
            at = self._last_occupied(idx)
            layer = self.layers[0 if at is None else at]
            x, z = PAULI_BITS[name]
            layer.paulis[idx[0]][0] = (layer.paulis[idx[0]][0] + x) % 2
            layer.paulis[idx[0]][1] = (layer.paulis[idx[0]][1] + z) % 2
        else:
            raise ValueError(f"Gate {gate} not implementable in MB GKP circuits.")

    # -- views ------------------------------------------------------------------------------------
    def layer_gates(self):
        """[[gate, ...], ...] per layer (``ClassicalControl`` wrappers included)."""
        return [list(layer.gates) for layer in self.layers]

    def describe(self):
        """JSON-able summary used by the golden tests: per layer the gate reprs and the
        recorded Pauli byproducts."""
        out = []
        for layer in self.layers:
            names = [repr(g.gate) + "?" if isinstance(g, ClassicalControl) else repr(g) for g in layer.gates]
            out.append({"gates": names, "paulis": [list(p) for p in layer.paulis]})
        return out


def noisy_circuit(circuit, noise, num_qubits: int | None = None) -> list:
    """The circuit re-ordered layer by layer with idle qubits padded by ``I`` and the
    GKP channel of ``noise`` after every gate *including the idle ones* -- the noise
    accounting of a measurement-based run, where every layer teleports every qubit.
    Pauli gates and classically controlled corrections are kept (noise-free for the
    Paulis, which are frame updates)."""
    lay = MBLayering(num_qubits if num_qubits is not None else max(max(g.indices) for g in circuit) + 1)
    paulis_after = {}                     # (layer index) -> Pauli gates recorded there, in program order
    for gate in circuit:
        name = type(gate).__name__
        if name in PAULI_BITS:
            at = lay._last_occupied(gate.indices)
            layer_index = 0 if at is None else len(lay.layers) + at
            paulis_after.setdefault(layer_index, []).append(gate)
        lay.add_gate(gate)
    lay.fill()
    out = []
    for li, layer in enumerate(lay.layers):
        for gate in layer.gates:
            if isinstance(gate, ClassicalControl):
                continue                   # MB-specific T correction: not part of the DV circuit
            out.append(gate)
            out.extend(noise.channels_after(gate))
        out.extend(paulis_after.get(li, []))
    return out
