"""Development tool: build the sharded schedule on the CPU (planner only, no state)
and print its structure.  usage: python tools/sched_probe.py [n] [ranks] [depth]"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from emu_backend import EmuBackend
from quantum_computations_b200 import sharded, workloads


def probe(n, ranks, depth, rank=0, opts=None, cls=None):
    be = EmuBackend()
    g = ranks.bit_length() - 1
    st = types.SimpleNamespace(n=n, g=g, n_local=n - g, backend=be, phys=list(range(n)), flip=[0] * g,
                               comm=types.SimpleNamespace(rank=rank, size=ranks))
    circ = workloads.sv_random_circuit(n, depth, 1234)
    sim = (cls or sharded.ShardedSimulator)(circ, st, plan_options=opts)
    t0 = time.perf_counter()
    sched = sim.compile()
    dt = time.perf_counter() - t0
    kinds = "".join("P" if it[0] == "plan" else "S" if it[0] == "exchange" else "X" for it in sched)
    return sim, sched, kinds, dt, len(circ)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 34
    ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    sim, sched, kinds, dt, ngates = probe(n, ranks, depth)
    print("gates", ngates, "compile s", round(dt, 2), sim.stats)
    print(kinds)
    print("passes per plan", [it[1].stats["n_passes"] for it in sched if it[0] == "plan"])
    print("qubits per exchange", [len(it[1]) for it in sched if it[0] == "exchange"])
