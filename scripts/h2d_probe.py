"""Diagnostic (GPU box): host->device bandwidth of page-locked buffers, one copy vs several concurrent streams, and
host-side packing bandwidth (numpy copy as a proxy for one thread's memcpy)."""
import time, torch, numpy as np
torch.cuda.init()
N = 13 << 20
tot = 24
src = torch.empty(tot * N, dtype=torch.uint8).pin_memory()
src.numpy()[:] = 1
dst = torch.empty(tot * N, dtype=torch.uint8, device="cuda")
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(tot):
            part = N // ns
            for j, st in enumerate(streams):
                with torch.cuda.stream(st):
                    dst[i * N + j * part:i * N + (j + 1) * part].copy_(src[i * N + j * part:i * N + (j + 1) * part], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D {ns} stream(s): {tot * N / dt / 1e9:.1f} GB/s")
# device reading pinned host memory directly (zero copy)
import ctypes
a = src.numpy()
t0 = time.perf_counter(); b = a[: 8 * N].copy(); dt = time.perf_counter() - t0
print(f"host single-thread copy: {8 * N / dt / 1e9:.1f} GB/s (read+write)")
import threading
def work(lo, hi, out): out.append(a[lo:hi].sum(dtype=np.uint64))
for nt in (1, 2, 4, 8):
    outs = []
    th = [threading.Thread(target=work, args=(i * (tot * N // nt), (i + 1) * (tot * N // nt), outs)) for i in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
    print(f"host read {nt} threads: {tot * N / dt / 1e9:.1f} GB/s")
