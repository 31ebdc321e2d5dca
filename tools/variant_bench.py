#!/usr/bin/env python
"""Development tool: time the bench circuit (config C4) under several planner option sets
in one process, kernels only (CUDA events around plan.execute).

    python tools/variant_bench.py --qubits 30 --variants '{}' '{"max_group":4}' ...

One JSON line per variant: passes, round trips, layers, ms per pass, gates/s, fraction of the
measured HBM copy peak per pass, final norm."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from quantum_computations_b200 import _capi, engine, workloads  # noqa: E402

if os.environ.get("QSIM_EXPERIMENT_LIB"):        # development builds of the library (other CTA sizes ...)
    engine._LIB_PATH = os.path.abspath(os.environ["QSIM_EXPERIMENT_LIB"])
from quantum_computations_b200.states import State  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--seed", type=int, default=30)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--variants", nargs="*", default=["{}"])
    args = ap.parse_args()
    n = args.qubits
    be = engine.get_backend()
    circuit = workloads.sv_random_circuit(n, args.depth, args.seed)
    ops = []
    for g in circuit:
        ops.extend(g.lowered(n, False))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    state = engine.DeviceState.product([State.ZERO.get()] * n, be)
    amps = np.ascontiguousarray(np.stack([np.asarray(State.ZERO.get(), dtype=np.complex128)] * n))

    def reset():
        _capi.check(be.lib, be.lib.qsim_init_product(be.ptr(state.buf), n,
                                                     amps.view(np.float64).ctypes.data_as(_capi.c_double_p),
                                                     be.stream()))

    for v in args.variants:
        opts = json.loads(v)
        env = opts.pop("env", None)
        plan = engine.Plan(be, n, ops, opts)
        reset()
        plan.execute(state.buf)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(args.reps):
            reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.execute(state.buf)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= args.reps
        passes = plan.stats["n_passes"]
        mspp = ms / passes
        print(json.dumps({"opts": opts, "passes": passes, "steps": plan.stats["n_steps"],
                          "layers": plan.stats["n_layers"], "ms": round(ms, 2), "ms_per_pass": round(mspp, 3),
                          "gates_per_s": round(len(circuit) / (ms * 1e-3), 1),
                          "frac": round(2 * 16 * 2.0 ** n / (mspp * 1e-3) / 1e9 / peak, 4),
                          "norm": state.norm()}), flush=True)


if __name__ == "__main__":
    main()
