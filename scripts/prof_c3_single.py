import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recfilter_b200 import Plan, Scan, gaussian_weights
G3 = gaussian_weights(5.0, 3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ext = (8192, 8192) if B == 1 else (8192, 8192, B)
plan = Plan(ext, "f32", [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)], "clamp")
src = torch.rand(B * 8192 * 8192, device="cuda"); dst = torch.empty_like(src)
for _ in range(3): plan.execute(src, dst)
torch.cuda.synchronize()
