#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (C1-C3), measured beside the CPU
oracle (the reference's algorithm).  Not the headline bench (that is bench.py);
prints one JSON line per configuration.

    python tools/extra_bench.py [--rb-sequences 10000] [--dm-qubits 12] [--dm-depth 100]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from golden.specs import as_oracle_ops, from_spec  # noqa: E402
from oracle import dense_ref, gkp_noise  # noqa: E402
from quantum_computations_b200 import channels, engine, gates, simulator, states, workloads  # noqa: E402
from quantum_computations_b200.batched import BatchedSimulator  # noqa: E402
from quantum_computations_b200.simulator import Simulator  # noqa: E402
from quantum_computations_b200.states import State  # noqa: E402


def rb(args):
    """C2: RB sequences, N=2, density matrix + GKP channel at 10 dB."""
    rng = np.random.default_rng(20251018)
    depths = [8, 10, 15, 20]
    t0 = time.perf_counter()
    circuits = [workloads.rb_random_circuit(2, depths[i % 4], rng) for i in range(args.rb_sequences)]
    gen_s = time.perf_counter() - t0
    ngates = sum(len(c) for c in circuits)
    noise = channels.GKPNoise(10.0)
    sim = BatchedSimulator(2, noise)
    sim.run(circuits[:64])                                   # warm-up (library load, opcode table)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sim.run(circuits)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    # CPU comparator: the reference's dense algorithm on a bounded sample of the same circuits
    sample = circuits[: args.rb_cpu_sample]
    t0 = time.perf_counter()
    worst = 0.0
    for i, circ in enumerate(sample):
        rho = np.zeros((4, 4), dtype=np.complex128)
        rho[0, 0] = 1.0
        psi = np.array([1, 0, 0, 0], dtype=np.complex128)
        for g in circ:
            rho = dense_ref.apply_matrix(rho, g.indices, g.matrix)
            psi = dense_ref.apply_matrix(psi, g.indices, g.matrix)
            for q, (px, pz) in zip(g.indices, noise.flips_for(g)):
                rho = dense_ref.apply_kraus(rho, [q], gkp_noise.pauli_flip_kraus(px, pz))
        worst = max(worst, abs(res["fidelity"][i] - dense_ref.fidelity(rho, psi)),
                    abs(res["purity"][i] - dense_ref.purity(rho)))
    cpu_s = time.perf_counter() - t0
    cpu_rate = len(sample) / cpu_s
    print(json.dumps({"config": "C2 RB", "sequences": args.rb_sequences, "gates": ngates,
                      "gpu_seconds_e2e": gpu_s, "sequences_per_s": args.rb_sequences / gpu_s,
                      "gates_per_s": ngates / gpu_s, "circuit_generation_seconds": gen_s,
                      "cpu_port_sequences_per_s": cpu_rate, "cpu_sample": len(sample), "cpu_cores": 1,
                      "speedup_vs_cpu_port": (args.rb_sequences / gpu_s) / cpu_rate,
                      "max_abs_diff_fidelity_purity": worst,
                      "mean_fidelity": float(res["fidelity"].mean()), "mean_purity": float(res["purity"].mean())}))


def grover(args):
    """C1: 3-qubit Grover (rewritten circuit) on rho with the GKP channel."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "noisy_grover.npz"))
    meta = json.loads(str(z["meta"]))
    out = []
    for rec in meta:
        circ = [from_spec(s, gates, simulator, z, channels, states) for s in rec["circuit"]]
        noisy = channels.GKPNoise(rec["db"]).noisy(circ)
        rho0 = np.zeros((8, 8), dtype=np.complex128)
        rho0[0, 0] = 1.0
        sim = Simulator(noisy)
        sim.run(rho0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = sim.run(rho0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        err = float(np.abs(got - z[rec["out"]]).max() / np.abs(z[rec["out"]]).max())
        out.append({"tagged": rec["tagged"], "db": rec["db"], "success": float(sum(got[t, t].real for t in rec["tagged"])),
                    "seconds": dt, "rel_err_vs_reference": err, "ops": len(noisy)})
    print(json.dumps({"config": "C1 Grover 3q + GKP noise", "points": out}))


def dm(args):
    """C3: N-qubit density matrix, layered Clifford+T circuit with per-gate channels."""
    n, depth = args.dm_qubits, args.dm_depth
    noise = channels.GKPNoise(10.0)
    layers = workloads.dm_random_layers(n, depth, 12)
    circ = noise.noisy([g for layer in layers for g in layer])
    n_unitary = sum(len(l) for l in layers)
    be = engine.get_backend()
    ops = []
    for g in circ:
        ops.extend(g.lowered(n, True))
    t0 = time.perf_counter()
    plan = engine.Plan(be, 2 * n, ops)
    plan_s = time.perf_counter() - t0
    state = engine.DeviceState.product([State.ZERO.get()] * (2 * n), be)
    state.ndim = 2
    plan.execute(state.buf)
    torch.cuda.synchronize()
    state = engine.DeviceState.product([State.ZERO.get()] * (2 * n), be)
    state.ndim = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.execute(state.buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    passes = plan.stats["n_passes"]
    bytes_pass = 2 * 16 * 4.0 ** n
    print(json.dumps({"config": f"C3 DM {n}q depth {depth}", "unitaries": n_unitary, "ops_with_channels": len(circ),
                      "plan": plan.stats, "plan_seconds": plan_s, "ms_total": ms, "ms_per_pass": ms / passes,
                      "us_per_gate_or_channel": 1e3 * ms / len(circ), "GBps": bytes_pass * passes / (ms * 1e-3) / 1e9,
                      "trace": state.trace().real, "purity": state.purity()}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rb-sequences", type=int, default=10000)
    ap.add_argument("--rb-cpu-sample", type=int, default=100)
    ap.add_argument("--dm-qubits", type=int, default=12)
    ap.add_argument("--dm-depth", type=int, default=100)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    for name, fn in (("grover", grover), ("rb", rb), ("dm", dm)):
        if not a.only or a.only == name:
            fn(a)
