"""Diagnostic (GPU box): what slows the host->device copy of the end-to-end leg down (54 GB/s alone, ~30 GB/s inside the
pipeline)? Copies alone / with a kernel running / with the host spinning on an event / with device->host traffic."""
import time, threading, torch
torch.cuda.init()
N = 13 << 20
tot = 20
src = torch.empty(tot * N, dtype=torch.uint8).pin_memory(); src.numpy()[:] = 1
dst = torch.empty(tot * N, dtype=torch.uint8, device="cuda")
back = torch.empty(2 << 20, dtype=torch.uint8).pin_memory()
dev_small = torch.empty(2 << 20, dtype=torch.uint8, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
s_copy, s_k, s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
big = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
def run(kernel=False, spin=False, d2h=False, membound=False):
    torch.cuda.synchronize()
    stop = False
    ev = torch.cuda.Event()
    t0 = time.perf_counter()
    for i in range(tot):
        with torch.cuda.stream(s_copy):
            dst[i * N:(i + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
            ev.record(s_copy)
        if kernel:
            with torch.cuda.stream(s_k):
                (a @ a) if not membound else big.fill_(i)
        if d2h:
            with torch.cuda.stream(s_d2h):
                back.copy_(dev_small, non_blocking=True)
        if spin:
            while not ev.query():
                pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return tot * N / dt / 1e9
for kw in ({}, {"kernel": True}, {"kernel": True, "membound": True}, {"spin": True}, {"d2h": True}, {"kernel": True, "membound": True, "spin": True, "d2h": True}):
    print(kw, f"{run(**kw):.1f} GB/s")
