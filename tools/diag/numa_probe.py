"""Diagnostic: host topology of the GPU box and pinned H2D bandwidth as a function of where the pinned pages live."""
import os, subprocess, time, glob
import torch
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    print(n, open(n + "/cpulist").read().strip())
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
try:
    import pynvml
    pynvml.nvmlInit()
    p = torch.cuda.get_device_properties(0)
    print("torch uuid", getattr(p, "uuid", None), "pci", getattr(p, "pci_bus_id", None))
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
    print("nvml cpu affinity of GPU 0:", cpus[:4], "...", cpus[-4:], len(cpus), "uuid", pynvml.nvmlDeviceGetUUID(h))
except Exception as e:
    print("nvml failed", e)
dev = torch.device("cuda:0")
dst = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
full = os.sched_getaffinity(0)
def bw(cpus, label):
    os.sched_setaffinity(0, cpus)
    src = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    os.sched_setaffinity(0, full)
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {20 * 64 / 1024 / (e0.elapsed_time(e1) * 1e-3):.1f} GiB/s")
bw(full, "unbound")
for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    s = open(n + "/cpulist").read().strip()
    cp = set()
    for part in s.split(","):
        a, _, b = part.partition("-")
        cp.update(range(int(a), int(b or a) + 1))
    cp &= full
    if cp:
        bw(cp, os.path.basename(n) + f" ({len(cp)} cpus)")
