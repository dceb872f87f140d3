"""Both sweeps in one launch (fused_stream_kernel) against the two-sweep kernels: device time per image, result
difference, and a sweep of its scheduling knobs (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan, gaussian_weights
G3 = gaussian_weights(5.0, 3)
g4 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
PEAK = 6549.1


def timed(plan, src, dst, iters=20):
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize()
    plan.check()
    return a.elapsed_time(b) / iters


def case(N, B, variants):
    shape = (B, N, N) if B > 1 else (N, N)
    ext = (N, N, B) if B > 1 else (N, N)
    src = torch.rand(shape, device="cuda"); dst = torch.empty_like(src)
    ref = None
    for name, env in variants:
        for k in ("RFB_STREAM", "RFB_STREAM_LAG_A", "RFB_STREAM_LAG_P", "RFB_STREAM_DBG"): os.environ.pop(k, None)
        os.environ.update(env)
        plan = Plan(ext, "f32", [Scan(*s) for s in g4], "clamp")
        ms = timed(plan, src, dst)
        if ref is None: ref = dst.clone(); diff = 0.0
        else: diff = float((dst - ref).abs().max() / ref.abs().max())
        print(f"N={N} B={B} {name:28s}: {ms*1e3/B:7.1f} us/image  {B*N*N/ms/1e6:7.1f} Gsamples/s  12B-frac {12*B*N*N/ms/1e6/PEAK:.3f}  "
              f"diff vs first {diff:.2e}  launches={plan.num_launches}", flush=True)
        plan.close()


base = [("two sweeps", {"RFB_STREAM": "0"}), ("one launch", {"RFB_STREAM": "1"})]
for N, B in [(1024, 1), (2048, 1), (4096, 1), (1024, 16), (2048, 4), (4096, 2), (8192, 1), (8192, 8)]:
    case(N, B, base)
