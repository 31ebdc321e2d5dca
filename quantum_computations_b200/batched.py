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

import itertools
import operator

import numpy as np

from . import _capi
from . import numpy_quantum as npq
from .states import State


def _full_unitary(gate, n: int) -> np.ndarray:
    m = np.asarray(gate.matrix, dtype=np.complex128)
    if len(gate.indices) == n and list(gate.indices) == list(range(n)):
        return m
    return np.asarray(npq.expand_gate(m, n, list(gate.indices)), dtype=np.complex128)


# gate classes of gates.py whose matrix is a function of (class, angle) alone; anything
# else -- user subclasses, Gate / SingleQubitGate / TwoQubitGate instances -- is keyed by the
# matrix itself
_ZOO = frozenset(("I", "H", "X", "Y", "Z", "P", "Pdg", "T", "Tdg", "RZ", "CZ", "CX", "CNOT", "SWAP"))
_sim_ids = itertools.count()


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
        # the opcode of a gate object is remembered on the object under a name private to this
        # simulator (opcodes depend on its noise model), so that a circuit seen before is encoded
        # by a C-level attribute sweep instead of a Python loop
        self._attr = f"_qsb200_opcode_{next(_sim_ids)}"
        self._get_code = operator.attrgetter(self._attr)
        self._tables = None                      # device copy of the opcode tables: (count, tensor)
        self._zoo_ref: dict = {}

    # -- opcode table ------------------------------------------------------------------
    def _key(self, gate):
        cls = type(gate)
        matrix = np.asarray(gate.matrix, dtype=np.complex128)
        if cls.__name__ in _ZOO and cls.__module__.endswith("gates"):
            # the zoo shortcut only holds while the instance still carries the class's matrix
            zkey = (cls, tuple(gate.indices), getattr(gate, "angle", None))
            ref = self._zoo_ref.get(zkey)
            if ref is None:
                ref = self._zoo_ref[zkey] = _zoo_matrix_bytes(cls, gate)
            if ref == matrix.tobytes():
                return (cls.__module__, cls.__name__) + zkey[1:]
        return ("matrix", tuple(gate.indices), matrix.shape, matrix.tobytes())

    def _opcode(self, gate) -> int:
        if gate.matrix is None or gate.matrix.ndim != 2 or gate.matrix.shape[0] != gate.matrix.shape[1]:
            raise NotImplementedError(f"{gate}: only matrix gates can be batched")
        key = self._key(gate)
        code = self._codes.get(key)
        if code is not None:
            return code
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

    def _tag(self, circuits) -> None:
        """Slow path, once per gate object: work out the opcode and remember it on the gate."""
        attr, opcode = self._attr, self._opcode
        for circ in circuits:
            for gate in circ:
                if getattr(gate, attr, None) is None:
                    code = opcode(gate)
                    try:
                        setattr(gate, attr, code)
                    except AttributeError:                # __slots__ class: stays on the slow path
                        pass

    def encode(self, circuits):
        """(opcodes uint16, offsets int64) for a list of circuits."""
        lengths = np.fromiter(map(len, circuits), dtype=np.int64, count=len(circuits))
        offsets = np.zeros(len(circuits) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        total = int(offsets[-1])
        flat = itertools.chain.from_iterable(circuits)
        try:
            codes = np.fromiter(map(self._get_code, flat), dtype=np.uint16, count=total)
        except AttributeError:
            self._tag(circuits)
            try:
                flat = itertools.chain.from_iterable(circuits)
                codes = np.fromiter(map(self._get_code, flat), dtype=np.uint16, count=total)
            except AttributeError:
                codes = np.fromiter((self._opcode(g) for c in circuits for g in c), dtype=np.uint16, count=total)
        return codes, offsets

    def forget(self, circuits) -> None:
        """Drop the opcodes remembered on the gate objects of ``circuits`` (call it after
        editing a gate's matrix in place)."""
        for circ in circuits:
            for gate in circ:
                if hasattr(gate, self._attr):
                    delattr(gate, self._attr)

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
        # opcode tables: on the device once, again only when new opcodes appeared
        nops = max(1, len(self._superops))
        if self._tables is None or self._tables[0] != len(self._superops):
            sup = np.stack(self._superops) if self._superops else np.zeros((1, d * d, d * d), np.complex128)
            uni = np.stack(self._unitaries) if self._unitaries else np.zeros((1, d, d), np.complex128)
            self._tables = (len(self._superops), be.upload(sup.view(np.float64).reshape(-1)),
                            be.upload(uni.view(np.float64).reshape(-1)))
        _, d_sup, d_uni = self._tables
        # per call: one host->device copy (offsets | rho0 | psi0 | opcodes) and one copy back
        ncode = max(1, len(codes))
        head = np.concatenate([offsets.view(np.float64), np.ascontiguousarray(rho0).view(np.float64).reshape(-1),
                               np.ascontiguousarray(psi0).view(np.float64)])
        packed = np.empty(head.size * 8 + 2 * ncode + 6, dtype=np.uint8)
        packed[:head.size * 8] = head.view(np.uint8)
        packed[head.size * 8:head.size * 8 + 2 * len(codes)] = codes.view(np.uint8)
        d_in = be.upload(packed)
        base = be.ptr(d_in)
        p_off = base
        p_rho0 = p_off + 8 * offsets.size
        p_psi0 = p_rho0 + 16 * d * d
        p_codes = p_psi0 + 16 * d
        d_out = be.zeros(2 * B + (2 * d * d * B if return_rho else 0))
        p_fid = be.ptr(d_out)
        p_pur = p_fid + 8 * B
        p_rho = p_pur + 8 * B if return_rho else None
        lib = be.lib
        _capi.check(lib, lib.qsim_rb_batch(n, B, p_codes, p_off, nops, be.ptr(d_sup), be.ptr(d_uni), p_rho0, p_psi0,
                                           p_fid, p_pur, p_rho, be.stream()))
        host = be.download(d_out)
        out = {"fidelity": host[:B], "purity": host[B:2 * B]}
        if return_rho:
            out["rho"] = host[2 * B:].view(np.complex128).reshape(B, d, d)
        return out


def _zoo_matrix_bytes(cls, gate) -> bytes:
    """Bytes of the matrix the class itself gives a fresh instance on the same indices
    (b"" if the class cannot be re-instantiated that way)."""
    try:
        angle = getattr(gate, "angle", None)
        fresh = cls(*gate.indices) if angle is None else cls(*gate.indices, angle)
        return np.asarray(fresh.matrix, dtype=np.complex128).tobytes()
    except Exception:
        return b""


def run_replicas(simulator: BatchedSimulator, circuits, initial_state=None, *, rank: int = 0, world: int = 1,
                 gather=None) -> dict:
    """Independent circuits shard trivially (BASELINE north star: no communication): rank r of
    ``world`` runs the contiguous slice r of the batch on its own GPU.  ``gather`` (e.g.
    ``lambda x: all_gather_object(x)`` returning the list over ranks) joins the per-rank
    results in batch order; without it the local slice is returned together with its bounds."""
    B = len(circuits)
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    local = simulator.run(circuits[lo:hi], initial_state)
    if gather is None:
        local["slice"] = (lo, hi)
        return local
    parts = gather({k: np.asarray(v) for k, v in local.items()})
    return {k: np.concatenate([p[k] for p in parts]) for k in ("fidelity", "purity")}
