"""Development probe: D2H bandwidth into pinned memory allocated with and without the
process bound to the GPU's NUMA node."""
import os, time, glob
import torch

dev = torch.device("cuda", 0)
props = torch.cuda.get_device_properties(0)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
node_file = f"/sys/bus/pci/devices/{bus}/numa_node"
node = int(open(node_file).read()) if os.path.exists(node_file) else -1
print("gpu pci", bus, "numa node", node, "nodes", [os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node*")])
print("affinity now", len(os.sched_getaffinity(0)), "cpus")


def cpus_of(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def bw(label, n=1 << 29):
    src = torch.empty(n, dtype=torch.complex128, device=dev)
    src.fill_(1)
    dst = torch.empty(n, dtype=torch.complex128, pin_memory=True)
    dst.fill_(0)                     # touch
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        t = time.perf_counter(); dst.copy_(src); torch.cuda.synchronize(); dt = time.perf_counter() - t
        best = max(best, 16 * n / dt / 1e9)
    t = time.perf_counter(); src.copy_(dst); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(label, "D2H GB/s", round(best, 1), "H2D GB/s", round(16 * n / dt / 1e9, 1))
    del dst, src


bw("default")
full = os.sched_getaffinity(0)
if node >= 0:
    local = cpus_of(node) & full
    if local:
        os.sched_setaffinity(0, local)
        bw(f"bound to node {node} ({len(local)} cpus)")
        os.sched_setaffinity(0, full)
    for other in range(8):
        if other != node and os.path.exists(f"/sys/devices/system/node/node{other}"):
            oc = cpus_of(other) & full
            if oc:
                os.sched_setaffinity(0, oc)
                bw(f"bound to node {other} ({len(oc)} cpus)")
                os.sched_setaffinity(0, full)
                break
