import time, torch, gc
dev = torch.device("cuda", 0)
torch.cuda.init()
def sync(): torch.cuda.synchronize()
for rep in range(6):
    sync(); t = time.perf_counter()
    x = torch.empty(1 << 30, dtype=torch.complex128, device=dev)
    sync(); a = time.perf_counter() - t
    t = time.perf_counter(); x.fill_(1); sync(); b = time.perf_counter() - t
    t = time.perf_counter(); del x; sync(); c = time.perf_counter() - t
    print(f"rep {rep}: alloc {a:.4f} fill {b:.4f} free {c:.4f} reserved {torch.cuda.memory_reserved() >> 30} GiB allocated {torch.cuda.memory_allocated() >> 30}")
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_computations_b200 import engine
from quantum_computations_b200.states import State
be = engine.get_backend(0)
import numpy as np
for rep in range(6):
    sync(); t = time.perf_counter()
    st = engine.DeviceState.product([State.ZERO.get()] * 30, be)
    sync(); a = time.perf_counter() - t
    t = time.perf_counter(); del st; sync(); c = time.perf_counter() - t
    print(f"product rep {rep}: {a:.4f} free {c:.4f} reserved {torch.cuda.memory_reserved() >> 30} GiB")
