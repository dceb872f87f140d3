"""Device timing of the one-way configs (C1, C2, C4): single-pass look-back kernels vs the two-sweep kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan

def timeit(plan, src, dst, iters=20):
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize()
    plan.check()
    return a.elapsed_time(b) / iters

def run(name, ext, dtype, scans, border="zero", ts=None, **kw):
    if ts: os.environ["RFB_LB_TS"] = str(ts)
    plan = Plan(ext, dtype, [Scan(*s) for s in scans], border, **kw)
    os.environ.pop("RFB_LB_TS", None)
    n = int(np.prod(ext))
    src = (torch.rand(n, device="cuda") if dtype == "f32" else torch.randint(0, 255, (n,), device="cuda", dtype=torch.int32))
    dst = torch.empty_like(src)
    ms = timeit(plan, src, dst)
    kind = "look-back" if "single-pass" in plan.describe() else "two-sweep"
    print(f"{name:34s} {kind:10s} ts={ts or 'auto':>4} {ms*1e3:9.1f} us  {n/ms/1e6:8.1f} Gsamples/s  {8*n/ms/1e6:8.1f} GB/s algorithmic "
          f"({8*n/ms/1e6/6549.1:.3f} of measured peak)  launches={plan.num_launches}", flush=True)
    plan.close()

SAT = [(0, True, [1, 1]), (1, True, [1, 1])]
A8 = [1.0] + [0.01] * 8
only = sys.argv[1] if len(sys.argv) > 1 else ""
if only in ("", "c1"):
    for ts in (64, 128):
        run("C1 sat u32 2048^2", (2048, 2048), "u32", SAT, ts=ts)
    run("C1 sat u32 2048^2", (2048, 2048), "u32", SAT, engine="twopass")
if only in ("", "c2"):
    for ts in (64, 128):
        run("C2 sat f32 4096^2", (4096, 4096), "f32", SAT, ts=ts)
    run("C2 sat f32 4096^2", (4096, 4096), "f32", SAT, engine="twopass")
    run("box_filter_3 integral 4096^2", (4096, 4096), "f32", [(0, True, [1, 2, -1]), (1, True, [1, 2, -1])])
    run("sat f32 8192^2", (8192, 8192), "f32", SAT, ts=128)
    run("sat f32 8192^2", (8192, 8192), "f32", SAT, ts=64)
if only in ("", "c4"):
    run("C4 audio 64x2^24 r8", (1 << 24, 64), "f32", [(0, True, A8)])
    run("C4 audio 64x2^24 r8", (1 << 24, 64), "f32", [(0, True, A8)], engine="twopass")
    run("audio 64x2^24 r3", (1 << 24, 64), "f32", [(0, True, [0.3, 0.5, 0.1, 0.05])])
    run("prefix sum u32 64x2^24", (1 << 24, 64), "u32", [(0, True, [1, 1])])
