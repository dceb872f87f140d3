"""PCIe probe: pinned H2D, D2H, and both at once (what bounds bench.py's e2e figure)."""
import torch, time
n = 268435456 // 4
h_in = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
d = [torch.empty(n, device="cuda") for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d[0].copy_(h_in[0], non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out[1].copy_(d[1], non_blocking=True)
def both():
    h2d(); d2h()
t = run(h2d); print(f"H2D 268 MB pinned: {t*1e3:.2f} ms  {0.268435456/t:.1f} GB/s")
t = run(d2h); print(f"D2H 268 MB pinned: {t*1e3:.2f} ms  {0.268435456/t:.1f} GB/s")
t = run(both); print(f"H2D + D2H concurrently: {t*1e3:.2f} ms  {2*0.268435456/t:.1f} GB/s total")
