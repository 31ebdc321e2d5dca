"""Batched executor for many small independent circuits (randomised benchmarking).

Replaces the per-sample Python loop of ``PAPER/randomised_benchmarking.py:65-75``
for the DV side: thousands of 1- or 2-qubit sequences are encoded as byte-code,
shipped to the GPU once, and every sequence is evolved by one thread -- the
density matrix (with an optional per-gate noise channel) and, next to it, the
ideal ket, so the fidelity ``<psi_ideal| rho |psi_ideal>`` and the purity
``tr rho^2`` the reference computes at randomised_benchmarking.py:72-73 come out
of the same kernel.  This workload is latency-bound, not bandwidth-bound:
256 bytes of state per sequence.

Each distinct (gate type, qubits, angle) becomes one opcode whose 4^n x 4^n
superoperator  S = S_noise . (U (x) conj U)  and n-qubit unitary U are built once
on the host.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from . import numpy_quantum as npq
from .states import State


def _full_unitary(gate, n: int) -> np.ndarray:
    m = np.asarray(gate.matrix, dtype=np.complex128)
    if len(gate.indices) == n and list(gate.indices) == list(range(n)):
        return m
    return np.asarray(npq.expand_gate(m, n, list(gate.indices)), dtype=np.complex128)


class BatchedSimulator:
    def __init__(self, num_qubits: int, noise=None, *, backend=None):
        if num_qubits not in (1, 2):
            raise NotImplementedError("BatchedSimulator handles registers of 1 or 2 qubits")
        self.num_qubits = num_qubits
        self.noise = noise
        self._backend = backend
        self._codes: dict = {}
        self._superops: list[np.ndarray] = []
        self._unitaries: list[np.ndarray] = []

    # -- opcode table ------------------------------------------------------------------
    def _opcode(self, gate) -> int:
        name = type(gate).__name__
        if name in ("Gate", "SingleQubitGate", "TwoQubitGate"):
            key = (name, tuple(gate.indices), np.asarray(gate.matrix, dtype=np.complex128).tobytes())
        else:
            key = (name, tuple(gate.indices), getattr(gate, "angle", None))
        code = self._codes.get(key)
        if code is not None:
            return code
        if gate.matrix is None or gate.matrix.shape[0] != gate.matrix.shape[1]:
            raise NotImplementedError(f"{gate}: only matrix gates can be batched")
        n = self.num_qubits
        u = _full_unitary(gate, n)
        sup = np.kron(u, np.conjugate(u))
        if self.noise is not None:
            for q, (px, pz) in zip(gate.indices, self.noise.flips_for(gate)):
                acc = np.zeros_like(sup)
                for k in self.noise.pauli_kraus(px, pz):
                    kf = np.asarray(npq.expand_gate(np.asarray(k, dtype=np.complex128), n, [q]))
                    acc += np.kron(kf, np.conjugate(kf))
                sup = acc @ sup
        code = len(self._superops)
        if code >= 65536:
            raise NotImplementedError("more than 65536 distinct gates in one batch")
        self._codes[key] = code
        self._superops.append(np.ascontiguousarray(sup))
        self._unitaries.append(np.ascontiguousarray(u))
        return code

    def encode(self, circuits):
        """(opcodes uint16, offsets int64) for a list of circuits."""
        lengths = np.fromiter((len(c) for c in circuits), dtype=np.int64, count=len(circuits))
        offsets = np.zeros(len(circuits) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        codes = np.empty(int(offsets[-1]), dtype=np.uint16)
        pos = 0
        opcode = self._opcode
        for circ in circuits:
            for gate in circ:
                codes[pos] = opcode(gate)
                pos += 1
        return codes, offsets

    # -- execution ---------------------------------------------------------------------------
    def run(self, circuits, initial_state=None, *, return_rho: bool = False) -> dict:
        """Evolve every circuit from the same initial state.  Returns
        ``{"fidelity": ndarray, "purity": ndarray[, "rho": ndarray (B, d, d)]}``."""
        from . import engine
        be = self._backend or engine.get_backend()
        n = self.num_qubits
        d = 2 ** n
        if initial_state is None:
            initial_state = [State.ZERO] * n
        if isinstance(initial_state, list):
            psi0 = npq.tensor(*(s.get() for s in initial_state)).astype(np.complex128)
        else:
            psi0 = np.asarray(initial_state, dtype=np.complex128)
        if psi0.shape != (d,):
            raise ValueError("initial state has the wrong size")
        rho0 = np.outer(psi0, np.conjugate(psi0))

        codes, offsets = self.encode(circuits)
        B = len(circuits)
        if B == 0:
            return {"fidelity": np.zeros(0), "purity": np.zeros(0)}
        if len(codes) == 0:
            codes = np.zeros(1, dtype=np.uint16)           # keep the device pointer valid
        sup = np.stack(self._superops) if self._superops else np.zeros((1, d * d, d * d), np.complex128)
        uni = np.stack(self._unitaries) if self._unitaries else np.zeros((1, d, d), np.complex128)

        d_codes, d_off = be.upload(codes), be.upload(offsets)
        d_sup, d_uni = be.upload(sup.view(np.float64)), be.upload(uni.view(np.float64))
        d_rho0 = be.upload(np.ascontiguousarray(rho0).view(np.float64))
        d_psi0 = be.upload(np.ascontiguousarray(psi0).view(np.float64))
        d_fid, d_pur = be.zeros(B), be.zeros(B)
        d_rho = be.zeros(B * 2 * d * d) if return_rho else None
        lib = be.lib
        _capi.check(lib, lib.qsim_rb_batch(
            n, B, be.ptr(d_codes), be.ptr(d_off), max(1, len(self._superops)), be.ptr(d_sup), be.ptr(d_uni),
            be.ptr(d_rho0), be.ptr(d_psi0), be.ptr(d_fid), be.ptr(d_pur),
            be.ptr(d_rho) if d_rho is not None else None, be.stream()))
        out = {"fidelity": be.download(d_fid), "purity": be.download(d_pur)}
        if return_rho:
            out["rho"] = be.download(d_rho).view(np.complex128).reshape(B, d, d)
        return out
