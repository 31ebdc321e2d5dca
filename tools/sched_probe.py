"""Development tool: build the sharded schedule of every rank on the CPU (planner only,
no state, ranks as threads) and print its structure.
usage: python tools/sched_probe.py [n] [ranks] [depth]"""
import os, sys, threading, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from emu_backend import EmuBackend
from quantum_computations_b200 import sharded, workloads


class _ThreadComm:
    """all-gather between the rank threads"""
    def __init__(self, rank, size, shared):
        self.rank, self.size, self._sh = rank, size, shared

    def allgather_object(self, obj):
        sh = self._sh
        sh["slots"][self.rank] = obj
        sh["barrier"].wait()
        out = list(sh["slots"])
        sh["barrier"].wait()
        return out


def probe(n, ranks, depth, opts=None, circuit=None):
    be = EmuBackend()
    g = ranks.bit_length() - 1
    circ = circuit or workloads.sv_random_circuit(n, depth, 1234)
    shared = {"slots": [None] * ranks, "barrier": threading.Barrier(ranks)}
    sims = [None] * ranks

    def work(rank):
        st = types.SimpleNamespace(n=n, g=g, n_local=n - g, backend=be, phys=list(range(n)), flip=[0] * g,
                                   comm=_ThreadComm(rank, ranks, shared))
        sim = sharded.ShardedSimulator(circ, st, plan_options=opts)
        sim.compile()
        sims[rank] = sim

    t0 = time.perf_counter()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return sims, time.perf_counter() - t0, len(circ)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 34
    ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    sims, dt, ngates = probe(n, ranks, depth)
    print("gates", ngates, "compile s (all ranks, threads)", round(dt, 2))
    for r, sim in enumerate(sims):
        print(r, sim.stats)
    sched = sims[0]._schedule
    print("passes per plan", [it[1].stats["n_passes"] for it in sched if it[0] == "plan"])
    print("dense per plan", [it[1].stats["n_dense"] for it in sched if it[0] == "plan"])
    print("qubits per exchange", [len(it[1]) for it in sched if it[0] == "exchange"])
