"""Run the C3 workload a few times (for ncu): python scripts/prof_c3.py [iters] [fuse_dims]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fuse = int(sys.argv[2]) if len(sys.argv) > 2 else -1
W = H = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
G3 = gaussian_weights(5.0, 3)
sc = [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)]
plan = Plan((W, H), "f32", sc, "clamp", fuse_dims=fuse)
print(plan.describe())
src = torch.rand(W * H, device="cuda")
dst = torch.empty_like(src)
for _ in range(iters):
    plan.execute(src, dst)
torch.cuda.synchronize()
print("done")
