#!/usr/bin/env python
"""Micro-benchmark of single tile passes (development tool, not the headline bench).

    python tools/passbench.py --qubits 28 --kind sign      # passes with no matrix step: pure load+store
    python tools/passbench.py --qubits 28 --kind dense --ops 8 --tile-bits 12

Prints ms per pass and the fraction of the measured HBM copy bandwidth.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from quantum_computations_b200 import engine, gates  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--kind", default="sign", choices=["sign", "dense"])
    ap.add_argument("--ops", type=int, default=8, help="dense 2x2 matrices per pass")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--max-group", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    n = args.qubits
    be = engine.get_backend()
    rng = np.random.default_rng(0)
    if args.kind == "sign":
        circ = [gates.CZ(0, 1), gates.CZ(2, n - 1), gates.Z(3)]
    else:
        # `ops` random unitaries on distinct high qubits (forces them into the tile)
        circ = []
        for q in range(args.ops):
            m = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
            circ.append(gates.Gate([q % (n - 6)], m))
    ops = []
    for g in circ:
        ops.extend(g.lowered(n, False))
    opts = {"tile_bits": args.tile_bits, "low_bits": args.low_bits, "max_group": args.max_group,
            "max_dense_ops": 64, "merge_1q": 2}
    plan = engine.Plan(be, n, ops, opts)
    state = engine.DeviceState.product([np.array([1.0, 0.0])] * n, be)
    for _ in range(3):
        plan.execute(state.buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        plan.execute(state.buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (args.reps * max(1, plan.stats["n_passes"]))
    gbs = 2 * 16 * 2.0 ** n / (ms * 1e-3) / 1e9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    print(json.dumps({"kind": args.kind, "n": n, "opts": opts, "plan": plan.stats, "ms_per_pass": ms,
                      "GBps": gbs, "frac": gbs / peak}))


if __name__ == "__main__":
    main()
