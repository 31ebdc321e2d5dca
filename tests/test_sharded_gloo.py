"""Multi-rank path on CPU: world_size 2 and 4 over the gloo backend, every rank
running the host emulator on its shard.  Checks the swap scheduling, the
pack/unpack kernels' index arithmetic, the diagonal-gate restriction to rank
bits and the final gather against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, depth, seed, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emu_backend import emu
        from golden.specs import as_oracle_ops
        from oracle import strided
        from parity_cases import random_circuit
        from quantum_computations_b200 import sharded, workloads
        from quantum_computations_b200.states import State

        be = emu()
        comm = sharded.Comm()
        rng = np.random.default_rng(seed)                 # same circuit on every rank
        circ = workloads.sv_random_circuit(n, depth, seed) + random_circuit(n, 25, rng)
        vecs = [State.PLUS.get(), State.T.get()] + [State.ZERO.get()] * (n - 2)
        st = sharded.ShardedState(n, comm, backend=be, as_torch=torch.from_numpy)
        st.set_product(vecs)
        sharded.ShardedState.CHUNK_LOG2 = 5                 # several pipeline chunks per swap even at this size
        sim = sharded.ShardedSimulator(circ, st, plan_options=dict(tile_bits=6, low_bits=2))
        sim.run()
        nrm = st.norm()
        got = st.gather_numpy()
        if rank == 0:
            psi0 = np.ones(1)
            for v in vecs:
                psi0 = np.kron(psi0, v)
            ref, _ = strided.run(as_oracle_ops(circ), psi0.astype(np.complex128))
            err = float(np.abs(got - ref).max() / np.abs(ref).max())
            np.save(os.path.join(out_dir, "result.npy"),
                    np.array([err, nrm, np.linalg.norm(ref), sim.stats["swaps"], comm.bytes_exchanged]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 9), (4, 10)])
def test_sharded_matches_oracle(tmp_path, world, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 6, 34, str(tmp_path)), nprocs=world, join=True)
    err, nrm, ref_norm, swaps, nbytes = np.load(tmp_path / "result.npy")
    assert err < 1e-12
    assert abs(nrm - ref_norm) < 1e-12
    assert swaps > 0 and nbytes > 0                      # the global-qubit path was exercised
