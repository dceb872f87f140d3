"""C3 timing aid: whole-filter device time (events around N back-to-back executes) + per-stage times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
G3 = gaussian_weights(5.0, 3)
sc = [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)]
plan = Plan((W, H), "f32", sc, "clamp")
srcs = [torch.rand(W * H, device="cuda") for _ in range(4)]
dst = torch.empty_like(srcs[0])
for s in srcs: plan.execute(s, dst)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    a.record()
    for i in range(20): plan.execute(srcs[i % 4], dst)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 20)
plan.stage_timing(True)
for i in range(8): plan.execute(srcs[i % 4], dst)
torch.cuda.synchronize()
st = plan.stage_times()
print(f"{os.environ.get('TAG','')} C3 {W}x{H}: {best*1e3:.1f} us/image  {W*H/best/1e6:.1f} Gsamples/s   stages(us, incl ~6us event overhead each):",
      {k: round(v['ms'] * 1e3 / 8, 1) for k, v in st.items() if v['launches']}, flush=True)
