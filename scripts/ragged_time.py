"""Timing aid: image sizes that are not multiples of the fused tile (generic engine) beside padded-size runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
G3 = gaussian_weights(5.0, 3)
sc = [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)]
for (W, H) in [(1920, 1080), (1920, 1088), (3840, 2160), (3840, 2176), (8192, 8192 - 8)]:
    plan = Plan((W, H), "f32", sc, "clamp")
    src = torch.rand(W * H, device="cuda"); dst = torch.empty_like(src)
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    kind = "fused" if "fused pass" in plan.describe() else "generic"
    print(f"{W}x{H}: {ms*1e3:8.1f} us  {W*H/ms/1e6:7.1f} Gsamples/s  ({kind} engine)", flush=True)
