#!/usr/bin/env python
"""Development tool: time dense k-qubit blocks (5 <= k <= 8) on an n-qubit ket (k_dense_block,
FP64 tensor-core MMA) and check them against the tile pass applying the same unitary as a
product of its single-qubit factors where that is possible (random product unitary)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from quantum_computations_b200 import engine, gates
from quantum_computations_b200.states import State


def haar(d, rng):
    q, r = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.qubits
    be = engine.get_backend()
    rng = np.random.default_rng(5)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for k in (5, 6, 7, 8):
        for where in ("high", "low"):
            qs = list(range(k)) if where == "high" else list(range(n - k, n))
            u = haar(2 ** k, rng)
            g = gates.Gate(qs, u)
            ops = g.lowered(n, False)
            plan = engine.Plan(be, n, ops, {})
            st = engine.DeviceState.product([np.array([0.6, 0.8j])] * n, be)
            plan.execute(st.buf)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                plan.execute(st.buf)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            flops = 8.0 * (2 ** k) * 2.0 ** n            # complex MAC = 8 real flops per matrix entry and amplitude
            print(json.dumps({"k": k, "qubits": where, "n": n, "plan": plan.stats["n_generic"], "ms": round(ms, 3),
                              "GBps": round(32 * 2.0 ** n / ms / 1e6, 1), "hbm_frac": round(32 * 2.0 ** n / ms / 1e6 / peak, 3),
                              "TFLOPs_fp64": round(flops / ms / 1e9, 2), "norm": st.norm()}), flush=True)


if __name__ == "__main__":
    main()
