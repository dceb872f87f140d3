"""Timing aid for the small configs (C1 SAT 2048^2 u32, C2 SAT 4096^2 f32): python scripts/small_time.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan
for name, n, dt, tdt in (("C1 u32 2048", 2048, "u32", torch.int32), ("C2 f32 4096", 4096, "f32", torch.float32)):
    plan = Plan((n, n), dt, [Scan(0, True, [1, 1]), Scan(1, True, [1, 1])])
    src = torch.rand(n * n, device="cuda") if dt == "f32" else torch.randint(0, 255, (n * n,), device="cuda", dtype=tdt)
    dst = torch.empty_like(src)
    for _ in range(5): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        a.record()
        for _ in range(50): plan.execute(src, dst)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 50)
    print(f"{os.environ.get('TAG','')} {name}: {best*1e3:.1f} us  {n*n/best/1e6:.1f} Gsamples/s  | {plan.describe().splitlines()[1].strip()[:70]}", flush=True)
