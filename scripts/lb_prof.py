"""One launch sequence of a look-back config for ncu (python scripts/lb_prof.py c4|c4small|c2|c1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan
which = sys.argv[1] if len(sys.argv) > 1 else "c4small"
eng = sys.argv[2] if len(sys.argv) > 2 else "auto"
SAT = [(0, True, [1, 1]), (1, True, [1, 1])]
cfg = {"c4": ((1 << 24, 64), "f32", [(0, True, [1.0] + [0.01] * 8)]),
       "c4small": ((1 << 22, 64), "f32", [(0, True, [1.0] + [0.01] * 8)]),
       "c2": ((4096, 4096), "f32", SAT), "c1": ((2048, 2048), "u32", SAT)}[which]
plan = Plan(cfg[0], cfg[1], [Scan(*s) for s in cfg[2]], engine=eng)
n = int(np.prod(cfg[0]))
src = torch.rand(n, device="cuda") if cfg[1] == "f32" else torch.randint(0, 255, (n,), device="cuda", dtype=torch.int32)
dst = torch.empty_like(src)
for _ in range(3): plan.execute(src, dst)
torch.cuda.synchronize(); plan.check()
print(plan.describe())
