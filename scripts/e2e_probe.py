"""e2e probe: rf_plan_execute_host_batch timing against the number of images."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
W = H = 8192
G3 = gaussian_weights(5.0, 3)
plan = Plan((W, H), "f32", [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)], "clamp")
nmax = 8
hin = [torch.rand((H, W)).pin_memory() for _ in range(nmax)]
hout = [torch.empty((H, W)).pin_memory() for _ in range(nmax)]
for n in (1, 2, 4, 8):
    ins, outs = [h.data_ptr() for h in hin[:n]], [h.data_ptr() for h in hout[:n]]
    plan.realize_batch_ptr(ins, outs)
    t0 = time.perf_counter()
    for _ in range(3): plan.realize_batch_ptr(ins, outs)
    t = (time.perf_counter() - t0) / 3
    print(f"batch of {n}: {t*1e3:.2f} ms  ({t*1e3/n:.2f} ms per image, {n*W*H/t/1e9:.1f} Gsamples/s)", flush=True)
