"""A/B timing of the headline filter with the package found under argv[1] (development aid): stack of 8 and one image."""
import sys, os
sys.path.insert(0, sys.argv[1])
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
import recfilter_b200
G3 = gaussian_weights(5.0, 3)
g4 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
N = 8192
for B in (8, 1):
    plan = Plan((N, N, B) if B > 1 else (N, N), "f32", [Scan(*s) for s in g4], "clamp")
    src = torch.rand(B * N * N, device="cuda"); dst = torch.empty_like(src)
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): plan.execute(src, dst)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    plan.stage_timing(True)
    for _ in range(5): plan.execute(src, dst)
    torch.cuda.synchronize()
    st = plan.stage_times(); plan.stage_timing(False)
    print(os.path.dirname(recfilter_b200.__file__)[-28:], f"B={B}: {best*1e3/B:7.1f} us/image",
          {k: round(v["ms"] * 1e3 / 5 / B, 1) for k, v in st.items() if v["launches"]}, flush=True)
    plan.close()
