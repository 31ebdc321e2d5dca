"""Kraus channels and the finite-squeezing GKP logical-noise model.

The reference's ``dv_simulator`` has no noise channel of its own (SURVEY.md
section 0.3): its only Kraus applier is ``quantum_channel`` in
``PAPER/tomography.py:14-41`` (rho -> sum_i K_i rho K_i^dagger with full-size
K_i), and its GKP noise comes out of a continuous-variable simulation.  This
module adds, behind the same ``Gate.apply`` protocol,

* ``Channel(indices, kraus, weights=None)``: a k-qubit Kraus channel acting on a
  density matrix.  On the device rho is its row-major vec, so the whole channel
  is ONE 4^k x 4^k matrix  sum_i w_i K_i (x) conj(K_i)  on qubits
  (indices, indices + N) -- a single fused-kernel op, however many K_i.
* ``GKPNoise(db)``: the closed-form per-gate logical error model the reference
  plots as its "Analytical estimate" (``PAPER/plot_data.ipynb:64-75`` with
  ``db2eps`` from ``GKP/utils.py:14-15``), turned into single-qubit Pauli
  channels: an X flip with probability ``px`` and an independent Z flip with
  probability ``pz`` after every gate, on each qubit the gate touches.
"""
from __future__ import annotations

import math

import numpy as np

from . import numpy_quantum as npq
from .gates import Gate


class Channel(Gate):
    """rho -> sum_i [w_i] K_i rho K_i^dagger on the listed qubits."""

    def __init__(self, indices: list[int], kraus, weights=None):
        super().__init__(list(indices), None)
        dim = 2 ** len(indices)
        ks = [np.asarray(k) for k in kraus]
        if not ks:
            raise ValueError("A channel needs at least one Kraus operator.")
        for k in ks:
            if k.shape != (dim, dim):
                raise ValueError("Dimensions of given matrix is not compatible with number of indices.")
        if weights is not None and len(weights) != len(ks):
            raise ValueError("One weight per Kraus operator required.")
        self.kraus = ks
        self.weights = None if weights is None else [float(w) for w in weights]

    def superoperator(self) -> np.ndarray:
        """sum_i w_i K_i (x) conj(K_i): acts on the row-major vec restricted to
        (row qubits, column qubits) of the channel's support."""
        dim = self.kraus[0].shape[0]
        sup = np.zeros((dim * dim, dim * dim), dtype=np.complex128)
        for i, k in enumerate(self.kraus):
            term = np.kron(k, np.conjugate(k))
            sup += term if self.weights is None else self.weights[i] * term
        return sup

    def lowered(self, num_qubits: int, is_density: bool):
        if not is_density:
            raise TypeError(f"{self}: a Kraus channel needs a density matrix, not a ket.")
        if max(self.indices) >= num_qubits:
            raise ValueError("new_ordering must be a permutation of all qubits")
        targets = list(self.indices) + [i + num_qubits for i in self.indices]
        return [(targets, self.superoperator())]

    def result_dtype(self, num_qubits: int, state_dtype) -> np.dtype:
        return np.dtype(np.complex128)

    @property
    def _fusable(self) -> bool:
        return True


def unitary_channel_superop(matrix: np.ndarray) -> np.ndarray:
    """U (x) conj(U): vec(U rho U^dagger) = (U (x) conj U) vec(rho), row-major."""
    return np.kron(matrix, np.conjugate(matrix))


# ---- GKP finite-squeezing noise ------------------------------------------------------------
def db2eps(db_squeezing: float) -> float:
    """eps = 2 atanh(10^(-dB/10) / 2)   (GKP/utils.py:14-15)."""
    return 2.0 * math.atanh(math.pow(10.0, -db_squeezing / 10.0) / 2.0)


def eps2db(epsilon: float) -> float:
    return -10.0 * math.log10(2.0 * math.tanh(epsilon / 2.0))


_I2 = np.identity(2)
_X2 = np.array([[0.0, 1.0], [1.0, 0.0]])
_Z2 = np.array([[1.0, 0.0], [0.0, -1.0]])

# gate type name -> per-qubit (k_x, k_z): the integer multiplying eps/2 in the
# variance of the quadrature whose sqrt(pi) slip flips X resp. Z.
# Teleportation through one cluster node adds two units to each quadrature; a
# shear (P-type and CZ) feeds a third unit into the momentum quadrature, whose
# slips are logical Z flips.  Paulis are frame updates and cost nothing.
_NOISE_TYPES = {
    "I": [(2, 2)], "H": [(2, 2)],
    "P": [(2, 3)], "Pdg": [(2, 3)], "T": [(2, 3)], "Tdg": [(2, 3)], "RZ": [(2, 3)],
    "CZ": [(2, 3), (2, 3)],
    "CX": [(2, 3), (3, 2)],
    "SWAP": [(2, 2), (2, 2)],
    "X": [], "Y": [], "Z": [],
}


class GKPNoise:
    """Per-gate logical Pauli noise of GKP qubits at ``db`` dB of squeezing."""

    def __init__(self, db: float):
        self.db = float(db)
        self.epsilon = db2eps(self.db)

    def flip_probability(self, k: int) -> float:
        """e(k) = 1 - erf( sqrt( pi / (8 k eps / 2) ) )   (plot_data.ipynb:64-68)."""
        variance = k * self.epsilon / 2.0
        return 1.0 - math.erf(math.sqrt(math.pi / (8.0 * variance)))

    def gate_error_I(self) -> float:
        e2 = self.flip_probability(2)
        return 1.0 - (1.0 - e2) * (1.0 - e2)

    def gate_error_P(self) -> float:
        return 1.0 - (1.0 - self.flip_probability(2)) * (1.0 - self.flip_probability(3))

    @staticmethod
    def pauli_kraus(px: float, pz: float):
        """X with probability px, then Z with probability pz (independent)."""
        return [
            math.sqrt((1 - px) * (1 - pz)) * _I2,
            math.sqrt(px * (1 - pz)) * _X2,
            math.sqrt((1 - px) * pz) * _Z2,
            math.sqrt(px * pz) * (_Z2 @ _X2),
        ]

    def flips_for(self, gate) -> list[tuple[float, float]]:
        """[(px, pz), ...] for each qubit of ``gate`` (in ``gate.indices`` order)."""
        kinds = _NOISE_TYPES.get(type(gate).__name__)
        if kinds is None:
            kinds = [(2, 2)] * len(gate.indices)      # unknown gate: one teleportation step per qubit
        return [(self.flip_probability(kx), self.flip_probability(kz)) for kx, kz in kinds]

    def channels_after(self, gate) -> list[Channel]:
        out = []
        for q, (px, pz) in zip(gate.indices, self.flips_for(gate)):
            out.append(Channel([q], self.pauli_kraus(px, pz)))
        return out

    def noisy(self, circuit) -> list:
        """The circuit with each gate followed by its noise channel(s)."""
        from .simulator import ClassicalControl
        out = []
        for gate in circuit:
            out.append(gate)
            controlled = isinstance(gate, ClassicalControl)
            inner = gate.gate if controlled else gate
            if getattr(inner, "matrix", None) is None or inner.matrix.shape[0] != inner.matrix.shape[1]:
                continue                                # M / Insert: no channel
            for ch in self.channels_after(inner):
                out.append(ClassicalControl(ch, gate._pos, gate._neg) if controlled else ch)
        return out
