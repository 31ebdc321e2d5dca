"""Multi-GPU arm of bench.py (torchrun, one rank per GPU, NCCL over NVLink).

Configuration C5 of BASELINE.json: the 34-qubit complex128 random circuit (depth 200,
10 200 gates) on a register sharded by its top log2(N) qubits over N = 2 / 4 / 8 GPUs
(128 / 64 / 32 GiB of state per GPU).  The total work is the same for every N: strong
scaling.  Non-diagonal gates on a rank qubit end a stage; between stages up to log2(N)
rank qubits trade places with local ones in one all-to-all exchange, which is ONE kernel per
rank over NVLink peer memory (quantum_computations_b200/sharded.py, qsim_exchange_p2p).

`value` is plain gates/s of the 34-qubit circuit.  (`config.value_30q_equivalent` rescales it
by 2^(n-30) -- a gate on n qubits touches 2^n amplitudes -- to compare with the N = 1 line,
which runs the 30-qubit circuit because 34 qubits do not fit one GPU.)
"""
from __future__ import annotations

import json
import time

import numpy as np


def run_sharded(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from bench import (PASS_PARAM_BYTES, ClockSampler, measured_peaks, metric_sharded, plan_options, secondary_rb)
    from quantum_computations_b200 import engine, sharded, workloads
    from quantum_computations_b200.states import State

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = world.bit_length() - 1
    n = args.qubits or 34
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
    state.collect_swap_time()
    torch.cuda.synchronize()
    dist.barrier()

    launches0 = engine.launch_count(backend)
    state.swap_seconds, state.swaps, state.amps_sent = 0.0, 0, 0
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
    state.collect_swap_time()
    t = torch.tensor([total_ms, state.swap_seconds], device=comm.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, swap_seconds = float(t[0].item()), float(t[1].item())
    launches = engine.launch_count(backend) - launches0
    norm = state.norm()
    swaps_done, amps_sent = state.swaps, state.amps_sent

    # correctness at full size, outside the timed region: a short circuit followed by its inverse
    # must bring |0...0> back (every stage, exchange and leftover of the sharded path is in it)
    check = None
    if not args.no_check:
        short = workloads.sv_random_circuit(n, args.check_depth, args.seed + 1)
        both = short + workloads.inverse_circuit(short)
        csim = sharded.ShardedSimulator(both, state, plan_options=opts)
        csim.compile()
        csim.prepare(zero)
        csim.run()
        # logical index 0: every logical bit 0 -> the rank whose bits equal the flip flags, local index 0
        owner = sum(f << i for i, f in enumerate(state.flip))
        amp0 = np.zeros(2)
        if rank == owner:
            v = state.buf[:1].cpu().numpy()[0]
            amp0[:] = [v.real, v.imag]
        amp0 = comm.allreduce_sum(amp0)
        overlap = float(amp0[0] ** 2 + amp0[1] ** 2)
        check = {"circuit": f"depth-{args.check_depth} random circuit and its inverse on {n} qubits "
                            f"({len(both)} gates, {csim.stats['swaps']} exchanges)",
                 "one_minus_overlap_with_zero_state": 1.0 - overlap, "norm": state.norm()}
        del csim

    # C2 next to it: independent RB sequences shard trivially over the ranks (replicas, no collective
    # on the data path; the results are gathered as Python objects)
    rb = None
    if not args.no_secondary:
        try:
            rb = secondary_rb(args.rb_sequences, args.cpu_seconds, rank=rank, world=world,
                              gather=comm.allgather_object)
        except Exception as exc:
            rb = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        passes = sim.stats["passes"]
        swaps = sim.stats["swaps"]
        ms_per_step = total_ms / args.steps
        swap_ms = 1e3 * swap_seconds / max(1, swaps_done)
        compute_ms = (ms_per_step - swap_ms * swaps) / max(1, passes)
        shard_bytes = 16.0 * 2.0 ** state.n_local
        achieved = 2.0 * shard_bytes / (compute_ms * 1e-3) / 1e9
        sent_bytes = 16.0 * amps_sent / max(1, swaps_done)             # mean per exchange and GPU, each way
        nvlink = sent_bytes / (swap_ms * 1e-3) / 1e9 if swaps else None
        raw = args.steps * ngates / (total_ms * 1e-3)
        line = {
            "metric": metric_sharded(n, args.depth), "value": raw, "unit": "gates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"C5: {n}-qubit complex128 random circuit, depth {args.depth}, {ngates} gates, "
                                   f"sharded over {world} GPUs by the top {g} qubits ({state.n_local} local qubits, "
                                   f"{shard_bytes / 2 ** 30:.0f} GiB per GPU); shards exceed L2, no flush needed",
                       "plan": sim.stats, "plan_options": opts, "plan_seconds": plan_seconds, "final_norm": norm,
                       "value_30q_equivalent": raw * 2.0 ** (n - 30), "amp_updates_per_s": raw * 2.0 ** n,
                       "note": "the N = 1 line of this bench runs the 30-qubit circuit (34 qubits do not fit one "
                               "GPU); value_30q_equivalent = value x 2^(n-30) is the comparable figure"},
            "roofline": {"bound": "hbm", "kernel": "k_tile_pass", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_launch": 2.0 * shard_bytes, "launches_per_step": passes,
                         "mean_launch_ms": compute_ms,
                         "note": "step time minus the exchanges, divided by the passes"},
            "roofline_nvlink": {"bound": "nvlink", "kernel": "k_exchange_p2p", "exchanges_per_step": swaps,
                                "bytes_each_way_per_gpu": sent_bytes, "shard_fraction_sent": sent_bytes / shard_bytes,
                                "mean_ms": swap_ms, "achieved": nvlink, "unit": "GB/s per direction per GPU",
                                "peak_measured_peer_copy": 770.0, "peak_nominal": 900.0,
                                "frac": (nvlink / 900.0) if nvlink else None,
                                "frac_of_measured": (nvlink / 770.0) if nvlink else None,
                                "share_of_step": swap_ms * swaps / ms_per_step,
                                "note": "one exchange = k rank qubits <-> k local qubits, all 2^k - 1 blocks in ONE "
                                        "kernel per rank over peer memory (1 - 2^-k of the shard leaves each GPU); "
                                        "CUDA events from behind the barrier that starts the exchange (all ranks have finished their "
                                        "passes) to behind the barrier that ends it, max over ranks"},
            "check": check, "secondary": [rb] if rb else [], "cpu_baseline": None,
            "e2e": {"value": raw, "unit": "gates/s",
                    "h2d_bytes_per_step": PASS_PARAM_BYTES * passes + 64 * n, "d2h_bytes_per_step": 16,
                    "state_resident": True,
                    "note": "same timed region (prepare + schedule through ShardedSimulator.run): 2^n amplitudes "
                            "(256 GiB) exceed host memory, so the state stays sharded on the GPUs and the host reads "
                            "back the norm only -- this is NOT a host-to-host figure"},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    state.close()
    dist.destroy_process_group()
