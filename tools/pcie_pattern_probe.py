"""What the upload of sihmpc_resident_step can reach on this box: the copy pattern of one e2e step (bench.py measure_e2e)
without any kernel -- per step x [B][22] and ysp [B][2] FP64 up from a 20-step pinned recording, du/J/status down.
usage: python tools/pcie_pattern_probe.py"""
import time, torch
B, nx, ny, nu, S = 131072, 22, 2, 2, 20
dev = torch.device("cuda", 0)
xs = torch.empty(S, B, nx, dtype=torch.float64).pin_memory(); xs.normal_()
ys = torch.empty(S, B, ny, dtype=torch.float64).pin_memory(); ys.normal_()
dx = torch.empty(B, nx, dtype=torch.float64, device=dev); dy = torch.empty(B, ny, dtype=torch.float64, device=dev)
ddu = torch.zeros(B, nu, dtype=torch.float64, device=dev); dJ = torch.zeros(B, dtype=torch.float64, device=dev)
dst = torch.zeros(B, dtype=torch.int32, device=dev)
hdu = torch.empty(B, nu, dtype=torch.float64).pin_memory(); hJ = torch.empty(B, dtype=torch.float64).pin_memory()
hst = torch.empty(B, dtype=torch.int32).pin_memory()
s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def step(i, nchunk, down, same):
    k = 0 if same else i % S
    per = B // nchunk
    for c in range(nchunk):
        lo, hi = c * per, (c + 1) * per
        with torch.cuda.stream(s_up):
            dx[lo:hi].copy_(xs[k, lo:hi], non_blocking=True)
            dy[lo:hi].copy_(ys[k, lo:hi], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(s_up)
        if down:
            with torch.cuda.stream(s_dn):
                s_dn.wait_event(ev)
                hdu[lo:hi].copy_(ddu[lo:hi], non_blocking=True)
                hJ[lo:hi].copy_(dJ[lo:hi], non_blocking=True)
                hst[lo:hi].copy_(dst[lo:hi], non_blocking=True)
    s_up.synchronize(); s_dn.synchronize()

for name, nchunk, down, same in (("1 chunk, up only", 1, False, False), ("4 chunks, up only", 4, False, False),
                                 ("4 chunks, up + down", 4, True, False), ("4 chunks, up + down, same 25 MB every step", 4, True, True),
                                 ("8 chunks, up + down", 8, True, False)):
    for i in range(5): step(i, nchunk, down, same)
    t0 = time.perf_counter()
    n = 40
    for i in range(n): step(i, nchunk, down, same)
    dt = (time.perf_counter() - t0) / n
    up = B * (nx + ny) * 8
    print("%-46s %.3f ms per step   upload %.1f GB/s   -> %.3g scenario-steps/s" % (name, dt * 1e3, up / dt / 1e9, B / dt))
