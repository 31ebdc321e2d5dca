"""Development probe: where the end-to-end time of Simulator.run goes (30 qubits)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quantum_computations_b200 import engine, workloads, simulator
from quantum_computations_b200.states import State

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
be = engine.get_backend(0)
circ = workloads.sv_random_circuit(n, 200, 30)
out = be.pinned_empty(1 << n)
sim = simulator.Simulator(circ)
init = [State.ZERO] * n

def sync():
    torch.cuda.synchronize()

for rep in range(4):
    sync(); t0 = time.perf_counter()
    t = time.perf_counter(); st = sim._initial(init); sync(); a = time.perf_counter() - t
    t = time.perf_counter()
    ops = []
    for g in circ:
        ops.extend(g.lowered(n, False))
    b = time.perf_counter() - t
    t = time.perf_counter(); sim._apply_cached(st, ops); c0 = time.perf_counter() - t; sync(); c = time.perf_counter() - t
    t = time.perf_counter(); st.to_numpy(out=out); sync(); d = time.perf_counter() - t
    t = time.perf_counter(); del st; sync(); e = time.perf_counter() - t
    print(f"rep {rep}: initial {a:.4f} lowering {b:.4f} plan/launch {c0:.4f} exec {c:.4f} download {d:.4f} free {e:.4f} total {time.perf_counter() - t0:.4f}")
for rep in range(3):
    sync(); t0 = time.perf_counter()
    sim.run(init, out=out)
    sync(); print("full run", round(time.perf_counter() - t0, 4))
