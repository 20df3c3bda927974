import os, time, torch
print("cpus allowed", len(os.sched_getaffinity(0)))
for n in sorted(os.listdir("/sys/devices/system/node")):
    if n.startswith("node"):
        print(n, open("/sys/devices/system/node/%s/cpulist" % n).read().strip())
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
bus = pynvml.nvmlDeviceGetPciInfo(h).busId
bus = bus.decode() if isinstance(bus, bytes) else bus
print("bus", bus)
p = "/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()
print("gpu numa", open(p).read().strip() if os.path.exists(p) else "n/a")
def bw(tag):
    n = 64 << 20
    hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
    hbuf.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(hbuf, non_blocking=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(hbuf, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(tag, "H2D GB/s %.1f" % (10 * n / e0.elapsed_time(e1) / 1e6))
torch.cuda.init()
bw("default affinity")
nodes = [n for n in sorted(os.listdir("/sys/devices/system/node")) if n.startswith("node")]
def parse(s):
    out = []
    for part in s.split(","):
        if "-" in part: a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part: out.append(int(part))
    return out
allowed = os.sched_getaffinity(0)
for n in nodes:
    cpus = set(parse(open("/sys/devices/system/node/%s/cpulist" % n).read().strip())) & allowed
    if not cpus: print(n, "no allowed cpus"); continue
    os.sched_setaffinity(0, cpus)
    bw("bound to " + n)
