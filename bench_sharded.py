"""Multi-GPU arm of bench.py (torchrun, one rank per GPU, NCCL over NVLink).

Configuration C5: the same random-circuit generator on a register sharded by its
top log2(N) qubits.  Default: 30 local qubits per GPU, i.e. 31 / 32 / 33 qubits on
2 / 4 / 8 GPUs -- exactly the per-GPU state of the N = 1 bench, so the runs form a
weak-scaling series; `--qubits 34` runs the 34-qubit case of BASELINE.json (31
local qubits on 8 GPUs; recorded in profiles/).  Non-diagonal gates on a rank qubit
end a stage; between stages up to log2(N) rank qubits trade places with local ones in one
all-to-all exchange (see quantum_computations_b200/sharded.py).

A gate on an n-qubit register touches 2^n amplitudes, so gates/s alone is not
comparable across register sizes: `value` is gates/s x 2^(n-30), the rate in units
of 30-qubit gate applications (identical to plain gates/s at N = 1); the raw figure
is kept in config.raw_gates_per_s.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run_sharded(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from bench import METRIC, ClockSampler, measured_peaks, plan_options
    from quantum_computations_b200 import engine, sharded, workloads
    from quantum_computations_b200.states import State

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = world.bit_length() - 1
    n = args.qubits or (30 + g)
    backend = engine.get_backend(local_rank)
    comm = sharded.Comm()
    comm.device = torch.device("cuda", local_rank)
    opts = plan_options(args)

    circuit = workloads.sv_random_circuit(n, args.depth, args.seed)
    ngates = len(circuit)
    state = sharded.ShardedState(n, comm, backend=backend)
    sim = sharded.ShardedSimulator(circuit, state, plan_options=opts)
    t0 = time.perf_counter()
    sim.compile()
    plan_seconds = time.perf_counter() - t0
    zero = [State.ZERO.get()] * n

    def step():
        sim.prepare(zero)
        sim.run()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()

    launches0 = engine.launch_count(backend)
    state.swap_seconds, state.swaps, state.amps_sent = 0.0, 0, 0
    bytes0 = comm.bytes_exchanged
    total_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            step()
            ev1.record()
            torch.cuda.synchronize()
            total_ms += ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], device=comm.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    launches = engine.launch_count(backend) - launches0
    norm = state.norm()

    if rank == 0:
        peak, peak_src = measured_peaks()
        passes = sim.stats["passes"]
        swaps = sim.stats["swaps"]
        ms_per_step = total_ms / args.steps
        swap_ms = 1e3 * state.swap_seconds / max(1, state.swaps)
        compute_ms = (ms_per_step - swap_ms * swaps) / max(1, passes)
        shard_bytes = 16.0 * 2.0 ** state.n_local
        achieved = 2.0 * shard_bytes / (compute_ms * 1e-3) / 1e9
        sent_bytes = 16.0 * getattr(state, "amps_sent", 0) / max(1, state.swaps)   # mean per exchange and GPU
        nvlink = sent_bytes / (swap_ms * 1e-3) / 1e9 if swaps else None
        raw = args.steps * ngates / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": raw * 2.0 ** (n - 30), "unit": "gates/s",
            "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"C5: {n}-qubit complex128 random circuit, depth {args.depth}, {ngates} gates, "
                                   f"sharded over {world} GPUs by the top {g} qubits ({state.n_local} local qubits, "
                                   f"{shard_bytes / 2 ** 30:.0f} GiB per GPU); shards exceed L2, no flush needed",
                       "plan": sim.stats, "plan_options": opts, "plan_seconds": plan_seconds, "final_norm": norm,
                       "value_definition": "gates/s x 2^(n-30): gate applications per second in units of a "
                                           "30-qubit register (equals plain gates/s at N=1); raw rate below",
                       "raw_gates_per_s": raw, "amp_updates_per_s": raw * 2.0 ** n},
            "roofline": {"bound": "hbm", "kernel": "k_tile_pass", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_launch": 2.0 * shard_bytes, "launches_per_step": passes,
                         "mean_launch_ms": compute_ms},
            "swap": {"count_per_step": swaps, "mean_ms": swap_ms, "bytes_each_way_per_gpu": sent_bytes,
                     "shard_fraction_sent": sent_bytes / shard_bytes,
                     "achieved_GBps_per_direction": nvlink, "peak_GBps_measured_peer_copy": 770.0,
                     "peak_GBps_nominal": 900.0, "frac_of_measured": (nvlink / 770.0) if nvlink else None,
                     "note": "one exchange = k rank qubits <-> k local qubits in one all-to-all (1 - 2^-k of the "
                             "shard leaves each GPU); host-timed with a device synchronize on both sides"},
            "cpu_baseline": None,
            "e2e": {"value": raw * 2.0 ** (n - 30), "unit": "gates/s",
                    "h2d_bytes_per_step": 25288 * passes + 64 * n, "d2h_bytes_per_step": 16,
                    "note": "same timed region: set_product + schedule execution through ShardedSimulator.run; "
                            "the state stays sharded on the GPUs (2^n amplitudes exceed host memory), the host reads "
                            "back the norm"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    dist.destroy_process_group()
