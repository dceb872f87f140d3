"""C5 (512^3, +-x +-y +-z, order 2): device time with the xy pass unsliced / cut into slices of whole z planes
(RFB_PIPE_SLICES: P1 of slice i+1 before P2 of slice i, so that pass 2 can find its planes in L2); per-stage times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan
W2 = [[1.0, 0.5, 0.25], [1.0, 0.5, 0.125], [1.0, 0.5, 0.0625], [1.0, 0.5, 0.125], [1.0, 0.5, 0.25], [1.0, 0.5, 0.0625]]
scans = [(0, True, W2[0]), (0, False, W2[1]), (1, True, W2[2]), (1, False, W2[3]), (2, True, W2[4]), (2, False, W2[5])]
n = 512 ** 3
src = torch.rand(n, device="cuda"); dst = torch.empty_like(src)
ref = None
for env in [{}, {"RFB_PIPE_SLICES": "4"}, {"RFB_PIPE_SLICES": "8"}, {"RFB_PIPE_SLICES": "16"}, {"RFB_PIPE_SLICES": "32"}, {"RFB_STREAM": "1"}]:
    for k in ("RFB_PIPE_SLICES", "RFB_STREAM"): os.environ.pop(k, None)
    os.environ.update(env)
    plan = Plan((512, 512, 512), "f32", [Scan(*s) for s in scans])
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize(); plan.check()
    ms = a.elapsed_time(b) / 10
    if ref is None: ref = dst.clone(); diff = 0.0
    else: diff = float((dst - ref).abs().max() / ref.abs().max())
    print(f"{str(env):32s} {ms*1e3:8.1f} us  {n/ms/1e6:7.1f} Gsamples/s  diff {diff:.1e}  launches={plan.num_launches}", flush=True)
    if not env: print(plan.describe())
    plan.close()
