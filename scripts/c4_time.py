"""C4 timing aid: order-8 causal IIR over 64 channels x 2^24 samples."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan
N, C = 1 << 24, 64
plan = Plan((N, C), "f32", [Scan(0, True, [1.0] + [0.01] * 8)])
print(plan.describe())
src = torch.rand(N * C, device="cuda") - 0.5
dst = torch.empty_like(src)
for _ in range(2): plan.execute(src, dst)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(5): plan.execute(src, dst)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
plan.stage_timing(True)
for i in range(3): plan.execute(src, dst)
torch.cuda.synchronize()
st = plan.stage_times()
print(f"C4: {ms*1e3:.1f} us  {N*C/ms/1e6:.1f} Gsamples/s  {8*N*C/ms/1e6:.0f} GB/s algorithmic; stages(us):",
      {k: round(v['ms'] * 1e3 / 3, 1) for k, v in st.items() if v['launches']}, "ws MB", plan.workspace_bytes / 1e6)
