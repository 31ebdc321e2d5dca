"""Pauli-trajectory (Monte-Carlo) execution of GKP-noisy circuits on kets.

SURVEY.md section 8f, rank 1: the reference's noisy runs track random Pauli
byproducts in a frame (``GKP/simulator.py:26-55``, ``commute``).  Here the same
mechanism is available at the DV level without a density matrix: every shot
samples, per gate and qubit, an X flip with probability ``px`` and a Z flip with
probability ``pz`` (``channels.GKPNoise``), inserts them as ordinary ``X`` / ``Z``
gates and runs the circuit on a ket -- O(2^N) per shot instead of O(4^N).  The
inserted Paulis are free on the GPU.  Small registers (<= 12 qubits, 1- and 2-qubit
gates) run ALL shots in one kernel launch -- ``qsim_traj_batch``: one CTA per shot, the
ket in shared memory, the shot's flips folded into the rows of each gate matrix, so the
circuit is encoded once and a shot costs what the noiseless circuit costs.  Larger
registers run shot by shot; there the planner's Pauli-frame pass slides every X / Z
through CZ / Z and merges it into the next non-diagonal gate of the qubit
(``csrc/planner.cpp``, ``merge_single_qubit``).  Averaging |psi><psi| over shots
converges to the density-matrix result of ``Simulator(noise.noisy(circuit))``.
"""
from __future__ import annotations

import numpy as np

from . import gates as _gates
from .simulator import ClassicalControl, Simulator


def sample_trajectory(circuit, noise, rng: np.random.Generator) -> list:
    """One noisy realisation of ``circuit``: after every matrix gate, an ``X`` with
    probability px and then a ``Z`` with probability pz on each of its qubits."""
    out = []
    for gate in circuit:
        out.append(gate)
        controlled = isinstance(gate, ClassicalControl)
        inner = gate.gate if controlled else gate
        if getattr(inner, "matrix", None) is None or inner.matrix.shape[0] != inner.matrix.shape[1]:
            continue
        for q, (px, pz) in zip(inner.indices, noise.flips_for(inner)):
            for cls, p in ((_gates.X, px), (_gates.Z, pz)):
                if rng.random() < p:
                    flip = cls(q)
                    out.append(ClassicalControl(flip, gate._pos, gate._neg) if controlled else flip)
    return out


def _batchable(circuit, n: int) -> bool:
    """Plain 1- and 2-qubit matrix gates on at most 12 qubits: the one-CTA-per-shot kernel."""
    if n < 1 or n > 12:
        return False
    for gate in circuit:
        m = getattr(gate, "matrix", None)
        if isinstance(gate, ClassicalControl) or m is None or m.ndim != 2 or m.shape[0] != m.shape[1]:
            return False
        if len(gate.indices) not in (1, 2) or m.shape[0] != 2 ** len(gate.indices):
            return False
    return True


def sample_flips(circuit, noise, shots: int, rng: np.random.Generator) -> np.ndarray:
    """(shots, F) uint8: for every shot, in circuit order, per gate and per qubit of the gate, the X
    flip and then the Z flip.  Consumes the generator exactly like ``shots`` successive calls of
    ``sample_trajectory`` (one uniform number per possible flip, drawn whether it fires or not)."""
    probs = []
    for gate in circuit:
        for px, pz in noise.flips_for(gate):
            probs += [px, pz]
    probs = np.asarray(probs, dtype=np.float64)
    return (rng.random((shots, probs.size)) < probs).astype(np.uint8)


def run_batch(circuit, flips: np.ndarray, initial_ket: np.ndarray, *, backend=None, observable=None,
              return_states: bool = False) -> dict:
    """All shots in ONE kernel launch (``qsim_traj_batch``: one CTA per shot, the ket in shared
    memory, the shot's Pauli flips folded into the rows of each gate matrix).  ``flips`` as from
    ``sample_flips``.  Returns the mean probabilities, per-shot fidelities with ``observable``
    and, on request, every shot's final ket."""
    from . import _capi, engine
    be = backend or engine.get_backend()
    psi0 = np.ascontiguousarray(initial_ket, dtype=np.complex128).reshape(-1)
    n = int(psi0.size).bit_length() - 1
    shots = int(flips.shape[0])
    ops = np.zeros((len(circuit), 4), dtype=np.int32)
    mats = []
    off = 0
    for i, gate in enumerate(circuit):
        k = len(gate.indices)
        bits = [n - 1 - int(q) for q in gate.indices]            # reference qubit q is index bit n-1-q
        ops[i] = (k, bits[0], bits[1] if k == 2 else 0, off)
        m = np.ascontiguousarray(gate.matrix, dtype=np.complex128).reshape(-1)
        mats.append(m)
        off += m.size
    mats = np.concatenate(mats) if mats else np.zeros(1, dtype=np.complex128)
    d_ops = be.upload(ops.reshape(-1)) if len(circuit) else be.upload(np.zeros(4, dtype=np.int32))
    d_mats = be.upload(mats.view(np.float64))
    d_flips = be.upload(np.ascontiguousarray(flips, dtype=np.uint8).reshape(-1) if flips.size else np.zeros(1, np.uint8))
    d_psi0 = be.upload(psi0.view(np.float64))
    d_obs = be.upload(np.ascontiguousarray(observable, dtype=np.complex128).view(np.float64)) \
        if observable is not None else None
    d_out = be.zeros(shots + psi0.size)
    d_states = be.zeros(2 * shots * psi0.size) if return_states else None
    p_fid = be.ptr(d_out)
    _capi.check(be.lib, be.lib.qsim_traj_batch(
        n, shots, len(circuit), be.ptr(d_ops), be.ptr(d_mats), be.ptr(d_flips), int(flips.shape[1]), be.ptr(d_psi0),
        be.ptr(d_obs) if d_obs is not None else None, p_fid if d_obs is not None else None, p_fid + 8 * shots,
        be.ptr(d_states) if d_states is not None else None, be.stream()))
    host = be.download(d_out)
    out = {"probabilities": host[shots:] / max(1, shots), "shots": shots}
    if observable is not None:
        out["fidelities"] = host[:shots]
        out["fidelity"] = float(host[:shots].mean()) if shots else 0.0
    if return_states:
        out["states"] = be.download(d_states).view(np.complex128).reshape(shots, psi0.size)
    return out


def run_trajectories(circuit, noise, initial_state, shots: int, rng=None, *, backend=None,
                     observable=None, batch: bool | None = None) -> dict:
    """Run ``shots`` noisy realisations from the same initial state.

    Returns the mean outcome probabilities (|amplitude|^2 averaged over shots) and,
    if ``observable`` (a ket) is given, the mean fidelity |<observable|psi>|^2 -- the
    trajectory estimate of <observable| rho |observable>.

    Circuits of 1- and 2-qubit matrix gates on at most 12 qubits run as one batched kernel
    launch (``run_batch``); anything else (measurements, classical control, wider gates, larger
    registers) runs shot by shot through ``Simulator``.  Both paths draw the same random numbers
    from ``rng``, so they sample the same trajectories."""
    from .simulator import parse_state
    rng = np.random.default_rng() if rng is None else rng
    psi0 = parse_state(initial_state)
    n = int(np.size(psi0)).bit_length() - 1
    if batch is None:
        batch = np.ndim(psi0) == 1 and _batchable(circuit, n)
    if batch:
        flips = sample_flips(circuit, noise, shots, rng)
        return run_batch(circuit, flips, psi0, backend=backend, observable=observable)
    probs = None
    fid = 0.0
    for _ in range(shots):
        psi = Simulator(sample_trajectory(circuit, noise, rng), backend=backend).run(initial_state)
        p = np.abs(psi) ** 2
        probs = p if probs is None else probs + p
        if observable is not None:
            fid += abs(np.vdot(observable, psi)) ** 2
    out = {"probabilities": probs / shots, "shots": shots}
    if observable is not None:
        out["fidelity"] = fid / shots
    return out
