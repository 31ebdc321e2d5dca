"""Development tool: compile the bench circuit with several planner options on the CPU
(planner only) and print passes / round trips / layers / matrices.
usage: python tools/plan_probe.py [n] [depth]"""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from emu_backend import EmuBackend
from quantum_computations_b200 import engine, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 200
be = EmuBackend()
circ = workloads.sv_random_circuit(n, depth, 30)
ops = [(list(g.indices), g.matrix) for g in circ]
for opts in [dict(), dict(max_layers=1), dict(max_layers=2), dict(max_layers=8), dict(max_group=4),
             dict(max_group=4, max_layers=8), dict(low_bits=3), dict(max_dense_ops=24), dict(max_dense_ops=28),
             dict(max_dense_ops=16), dict(max_dense_ops=12), dict(max_dense_ops=28, max_group=4),
             dict(low_bits=3, max_dense_ops=24), dict(tile_bits=13, max_dense_ops=28), dict(tile_bits=13, max_dense_ops=28, max_group=4)]:
    t0 = time.perf_counter()
    plan = engine.Plan(be, n, ops, opts)
    dt = time.perf_counter() - t0
    s = plan.stats
    print(opts, "passes", s["n_passes"], "steps", s["n_steps"], "layers", s["n_layers"], "dense", s["n_dense"],
          "sign", s["n_sign"], "warp_syncs", s["n_warp_syncs"], f"steps/pass {s['n_steps']/s['n_passes']:.2f}",
          f"compile {dt:.2f}s")
