"""Long 1-D signals with filter orders above 8 (apps/audio/audio_filter_high_order.cpp sweeps 1..29): device time and
per-stage times of the generic engine (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan

for (n, rows) in [(1 << 24, 1), (1 << 20, 64)]:
    for order in (7, 9, 15, 17, 29):
        plan = Plan((n, rows) if rows > 1 else (n,), "f32", [Scan(0, True, [1.0] + [0.01] * order)])
        src = torch.rand(n * rows, device="cuda"); dst = torch.empty_like(src)
        for _ in range(2): plan.execute(src, dst)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): plan.execute(src, dst)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        plan.stage_timing(True)
        for _ in range(2): plan.execute(src, dst)
        torch.cuda.synchronize()
        st = plan.stage_times()
        plan.stage_timing(False)
        print(f"{rows} x {n} order {order}: {ms:9.3f} ms  {n*rows/ms/1e6:8.2f} Gsamples/s  launches={plan.num_launches}  stages(us):",
              {k: round(v['ms'] * 1e3 / 2, 1) for k, v in st.items() if v['launches']}, flush=True)
        if order in (7, 9): print(plan.describe())
        plan.close()
