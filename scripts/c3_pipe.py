"""C3 stack of B images: device time per image with the stack pipelined in 1 / 2 / B slices (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan, gaussian_weights
G3 = gaussian_weights(5.0, 3)
g4 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = 8192
src = torch.rand((B, N, N), device="cuda"); dst = torch.empty_like(src)
ref = None
for slices in [1, 2, B] + ([8] if B > 8 else []):
    os.environ["RFB_PIPE_SLICES"] = str(slices)
    plan = Plan((N, N, B), "f32", [Scan(*s) for s in g4], "clamp")
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    if ref is None: ref = dst.clone(); same = True
    else: same = bool(torch.equal(ref, dst))
    print(f"B={B} slices={slices}: {ms*1e3/B:7.1f} us/image  {B*N*N/ms/1e6:7.1f} Gsamples/s  frac {8*B*N*N/ms/1e6/6549.1:.3f}  identical={same}  launches={plan.num_launches}", flush=True)
    plan.close()
