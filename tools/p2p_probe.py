#!/usr/bin/env python
"""Peer-copy bandwidth probe between GPU 0 and GPU 1 (one process, torch copies):
unidirectional and simultaneous bidirectional, copy-engine path."""
import json
import torch

assert torch.cuda.device_count() >= 2
n = 1 << 30
a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
a1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0, s1 = torch.cuda.Stream(device=0), torch.cuda.Stream(device=1)


def sync():
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)


def run(bidir, reps=8, pull=False):
    sync()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        e0.record(s0)
    for _ in range(reps):
        with torch.cuda.device(0), torch.cuda.stream(s0):
            if pull:
                b0.copy_(a1, non_blocking=True)      # issued on GPU 0: pull from GPU 1
            else:
                a1.copy_(a0, non_blocking=True)      # issued on GPU 0: push to GPU 1
        if bidir:
            with torch.cuda.device(1), torch.cuda.stream(s1):
                if pull:
                    b1.copy_(a0, non_blocking=True)
                else:
                    b0.copy_(b1, non_blocking=True)
    with torch.cuda.device(0):
        e1.record(s0)
    sync()
    ms = e0.elapsed_time(e1)
    return reps * n / (ms * 1e-3) / 1e9


run(False)
out = {"uni_push_GBps": run(False), "uni_pull_GBps": run(False, pull=True),
       "bidir_push_GBps_per_dir": run(True), "bidir_pull_GBps_per_dir": run(True, pull=True)}
print(json.dumps(out))
