#!/usr/bin/env python
"""Small run of the fused path for compute-sanitizer (racecheck / memcheck):
a 15-qubit layered circuit (tile = 2^12, several tiles, warp-synchronised runs,
sign blocks, rotation form) checked against the strided oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from golden.specs import as_oracle_ops  # noqa: E402
from oracle import strided  # noqa: E402
from quantum_computations_b200 import workloads  # noqa: E402
from quantum_computations_b200.simulator import Simulator  # noqa: E402
from quantum_computations_b200.states import State  # noqa: E402

n = 15
circ = workloads.sv_random_circuit(n, 6, 30)
sim = Simulator(circ)
got = sim.run([State.ZERO] * n)
psi0 = np.zeros(2 ** n, dtype=np.complex128)
psi0[0] = 1
ref, _ = strided.run(as_oracle_ops(circ), psi0)
err = float(np.abs(got - ref).max() / np.abs(ref).max())
print("plan", sim.last_stats, "rel err", err)
assert err < 1e-12
