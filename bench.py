#!/usr/bin/env python
"""Benchmark of the gate-application hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo, on GPUs
    python bench.py --impl reference [--steps K] [--warmup W]    # reference CPU algorithm

Workload (N=1): config C4 -- the 30-qubit complex128 random circuit, depth 200
(200 x (30 single-qubit gates from {H,T,RZ,X,P} + 15 CZ on a random perfect
matching) = 9000 gates, ``workloads.sv_random_circuit(30, 200, seed=30)``).
One *step* = the whole circuit applied to |0...0>.

Reported on one JSON line:
  value      gates/s with the state resident in HBM and the fused plan compiled
             (the timed region is the tile-pass kernels only, CUDA events).
  e2e        gates/s through the public API ``Simulator(circuit).run([ZERO]*n, out=pinned)``:
             host circuit objects in; lowering, plan lookup (the plan compiled by the
             warm-up run is reused through the simulator's content-keyed cache, compile
             time is config.plan_seconds), product state, launches and the device->host
             copy of the final 2^n amplitudes inside the timed region.
  roofline   k_tile_pass: algorithmic bytes per launch = 2 x 16 B x 2^n (one read and
             one write of the state) over the mean launch time, against the measured
             HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own algorithm (oracle/dense_ref.py: dense 2^N x 2^N
             operator per gate) timed on the host cores on a bounded sample of the
             same generator at the largest N that fits the time budget.

The state (16 GiB at n=30) is far larger than the 126 MB L2, so every pass
streams from HBM; no L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gates/sec (30q c128 random circuit, depth 200)"
# dram__bytes_read.sum + dram__bytes_write.sum per k_tile_pass launch at n = 30, default plan options, from the
# `ncu --set full` capture summarised in profiles/r1_ncu_full_k_tile_pass_n30.csv (17.18 GB + 17.12 GB)
NCU_TRAFFIC_N30 = 34.35e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="0 = 30 on one GPU, 30+log2(N) sharded")
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--seed", type=int, default=30)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--max-group", type=int, default=0)
    ap.add_argument("--max-dense", type=int, default=0)
    ap.add_argument("--lookahead", type=int, default=0)
    ap.add_argument("--max-layers", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per k_tile_pass launch from an ncu --set full capture")
    return ap.parse_args()


# ---- helpers ---------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_baseline(seconds: float, seed: int) -> dict:
    """Time the reference's algorithm (dense operator per gate, oracle/dense_ref.py)
    on the same generator at N=12, for about ``seconds`` of CPU work."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden.specs import as_oracle_ops
    from oracle import dense_ref
    from quantum_computations_b200 import workloads
    n = 12
    ops = as_oracle_ops(workloads.sv_random_circuit(n, 200, seed))
    state = np.zeros(2 ** n, dtype=np.complex128)
    state[0] = 1.0
    done, t0 = 0, time.perf_counter()
    while done < len(ops):
        state, _ = dense_ref.run([ops[done]], state)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"first {done} gates of the same generator at N={n} (the reference's dense 2^N x 2^N "
                      f"operator path cannot hold N=30); numpy {np.__version__}",
            "seconds": dt}


def plan_options(args) -> dict:
    return {"tile_bits": args.tile_bits, "low_bits": args.low_bits, "max_group": args.max_group,
            "max_dense_ops": args.max_dense, "lookahead": args.lookahead, "max_layers": args.max_layers}


# ---- reference arm -----------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = None
    for _ in range(max(1, args.warmup) - 1):
        cpu_baseline(min(2.0, args.cpu_seconds), args.seed)
    vals = []
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        base = cpu_baseline(args.cpu_seconds, args.seed)
        vals.append(base["value"])
    total = time.perf_counter() - t_all
    value = float(np.mean(vals))
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": {"workload": "C4 generator (sv_random_circuit, depth 200, seed 30) run by the reference's "
                                   "dense-operator algorithm at N=12; it cannot hold N=30"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- B200 arm ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from quantum_computations_b200 import engine, workloads
    from quantum_computations_b200.simulator import Simulator
    from quantum_computations_b200.states import State

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from bench_sharded import run_sharded
        return run_sharded(args, world, rank, local_rank)

    n = args.qubits or 30
    backend = engine.get_backend(local_rank)
    opts = plan_options(args)
    circuit = workloads.sv_random_circuit(n, args.depth, args.seed)
    ngates = len(circuit)

    # compile once (resident plan) for the kernel-only number
    t0 = time.perf_counter()
    ops = []
    for g in circuit:
        ops.extend(g.lowered(n, False))
    plan = engine.Plan(backend, n, ops, opts)
    plan_seconds = time.perf_counter() - t0
    stats = plan.stats
    passes = stats["n_passes"]

    state = engine.DeviceState.product([State.ZERO.get()] * n, backend)
    zero_amps = [State.ZERO.get()] * n

    def reset():
        lib = backend.lib
        amps = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.complex128) for a in zero_amps]))
        from quantum_computations_b200 import _capi
        _capi.check(lib, lib.qsim_init_product(backend.ptr(state.buf), n,
                                               amps.view(np.float64).ctypes.data_as(_capi.c_double_p),
                                               backend.stream()))

    for _ in range(args.warmup):
        reset()
        plan.execute(state.buf)
    torch.cuda.synchronize()

    launches0 = engine.launch_count(backend)
    kernel_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            reset()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            plan.execute(state.buf)
            ev1.record()
            torch.cuda.synchronize()
            kernel_ms += ev0.elapsed_time(ev1)
    launches = engine.launch_count(backend) - launches0 - args.steps      # minus the reset kernels
    norm = state.norm()
    value = args.steps * ngates / (kernel_ms * 1e-3)
    ms_per_step = kernel_ms / args.steps

    peak, peak_src = measured_peaks()
    bytes_per_launch = 2.0 * 16.0 * (2.0 ** n)
    launch_ms = kernel_ms / max(1, args.steps * passes)
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = args.traffic_bytes
    if traffic is None and n == 30 and not any(opts.values()):
        traffic = NCU_TRAFFIC_N30
    roofline = {"bound": "hbm", "kernel": "k_tile_pass", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "launches_per_step": passes,
                "mean_launch_ms": launch_ms}

    # end to end through the public API, host buffers on both sides
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 16 << n
        if psutil.virtual_memory().available > need * 1.5:
            out = backend.pinned_empty(1 << n)
            init = [State.ZERO] * n
            del state
            torch.cuda.empty_cache()
            sim = Simulator(circuit, plan_options=opts)
            sim.run(init, out=out)                              # warm-up
            e2e_steps = max(1, min(args.steps, 2))
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                sim.run(init, out=out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            assert abs(np.vdot(out[:1 << 20], out[:1 << 20]).real) >= 0.0
            h2d = passes * 25288 + 64 * n                       # kernel-parameter blocks + product-state amplitudes
            e2e = {"value": ngates / dt, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": need,
                   "seconds_per_step": dt, "steps": e2e_steps,
                   "note": "Simulator(circuit).run([ZERO]*n, out=pinned): lowering, plan lookup (the plan compiled "
                           "by the warm-up run is reused through the simulator's content-keyed cache; compiling takes "
                           "config.plan_seconds), product state, all passes, 2^n x 16 B device-to-host copy"}
        else:
            e2e = {"value": None, "unit": "gates/s", "skipped": "host memory too small for the 2^n output buffer"}

    base = None
    if not args.no_cpu_baseline:
        base = cpu_baseline(args.cpu_seconds, args.seed)

    line = {"metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"C4: {n}-qubit complex128 random circuit, depth {args.depth}, "
                                   f"{ngates} gates (sv_random_circuit seed {args.seed}); state 2^{n} x 16 B "
                                   "exceeds L2, no flush needed",
                       "plan": stats, "plan_options": opts, "plan_seconds": plan_seconds,
                       "gates_per_pass": ngates / max(1, passes), "final_norm": norm},
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary()}
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    main()
