"""Pauli-trajectory (Monte-Carlo) execution of GKP-noisy circuits on kets.

SURVEY.md section 8f, rank 1: the reference's noisy runs track random Pauli
byproducts in a frame (``GKP/simulator.py:26-55``, ``commute``).  Here the same
mechanism is available at the DV level without a density matrix: every shot
samples, per gate and qubit, an X flip with probability ``px`` and a Z flip with
probability ``pz`` (``channels.GKPNoise``), inserts them as ordinary ``X`` / ``Z``
gates and runs the circuit on a ket -- O(2^N) per shot instead of O(4^N).  The
inserted Paulis are free on the GPU: the planner's Pauli-frame pass slides every
X / Z through CZ / Z and merges it into the next non-diagonal gate of the qubit
(``csrc/planner.cpp``, ``merge_single_qubit``), so a noisy shot costs what the
noiseless circuit costs.  Averaging |psi><psi| over shots converges to the
density-matrix result of ``Simulator(noise.noisy(circuit))``.
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


def run_trajectories(circuit, noise, initial_state, shots: int, rng=None, *, backend=None,
                     observable=None) -> dict:
    """Run ``shots`` noisy realisations from the same initial state.

    Returns the mean outcome probabilities (|amplitude|^2 averaged over shots) and,
    if ``observable`` (a ket) is given, the mean fidelity |<observable|psi>|^2 -- the
    trajectory estimate of <observable| rho |observable>."""
    rng = np.random.default_rng() if rng is None else rng
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
