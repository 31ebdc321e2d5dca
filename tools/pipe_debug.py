#!/usr/bin/env python
"""Development tool: circuit-then-inverse at several sizes, reports the overlap with |0..0>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from quantum_computations_b200 import engine, workloads
from quantum_computations_b200.states import State

be = engine.get_backend()
for n in [int(a) for a in sys.argv[1:]] or [22, 23, 24, 25, 26]:
    depth = 6
    circ = workloads.sv_random_circuit(n, depth, int(os.environ.get("SEED", "30")))
    inv = workloads.inverse_circuit(circ) if os.environ.get("INV", "1") == "1" else []
    ops = []
    for g in circ + inv:
        ops.extend(g.lowered(n, False))
    plan = engine.Plan(be, n, ops, {})
    st = engine.DeviceState.product([State.ZERO.get()] * n, be)
    try:
        plan.execute(st.buf)
        torch.cuda.synchronize()
        amp0 = st.buf[0].item()
        print(n, plan.stats["n_passes"], "passes; <0|psi> =", amp0, "norm", st.norm(), flush=True)
    except Exception as exc:
        print(n, "FAILED", repr(exc)[:200], flush=True)
        break
