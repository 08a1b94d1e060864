"""Diagnostic (GPU box): what makes page-locked staging memory "DMA-warm"? Passes over a 252 MB page-locked buffer in
12.6 MB copies; then the CPU rewrites the buffer (as a producer filling its staging ring would) and the passes repeat."""
import time, torch, numpy as np
torch.cuda.init()
N = 4096 * 64 * 48
tot = 20
src = torch.empty(tot * N, dtype=torch.uint8).pin_memory()
dst = torch.empty(3 * N, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def one_pass(tag):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for i in range(tot):
            dst[(i % 3) * N:(i % 3 + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    print(f"{tag}: {tot * N / e0.elapsed_time(e1) / 1e6:.1f} GB/s", flush=True)
a = src.numpy()
a[:] = 1
for k in range(4):
    one_pass(f"after the first CPU fill, pass {k}")
t0 = time.perf_counter(); a[:] = 2; dt = time.perf_counter() - t0
print(f"CPU rewrite of the whole buffer: {a.nbytes / dt / 1e9:.1f} GB/s")
for k in range(3):
    one_pass(f"after a CPU rewrite, pass {k}")
# rewrite with a copy from pageable memory (what a producer does), a slice at a time, copy right behind it
pg = np.random.randint(0, 255, size=N, dtype=np.uint8)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
for i in range(tot):
    a[i * N:(i + 1) * N] = pg
    with torch.cuda.stream(s):
        dst[(i % 3) * N:(i % 3 + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"producer loop (CPU fill of a slice, then its copy): {tot * N / dt / 1e9:.1f} GB/s overall")
for k in range(2):
    one_pass(f"after the producer loop, pass {k}")
