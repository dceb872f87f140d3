"""Quick device timing of the BASELINE configs (development aid, not the bench contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recfilter_b200 import Plan, Scan, gaussian_weights

def timeit(plan, src, dst, iters=10):
    for _ in range(3): plan.execute(src, dst)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): plan.execute(src, dst)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def run(name, ext, dtype, scans, border="zero", **kw):
    plan = Plan(ext, dtype, [Scan(*s) for s in scans], border, **kw)
    n = int(np.prod(ext))
    tdt = {"f32": torch.float32, "u32": torch.int32}[dtype]
    src = (torch.rand(n, device="cuda") if dtype == "f32" else torch.randint(0, 255, (n,), device="cuda", dtype=tdt))
    dst = torch.empty_like(src)
    ms = timeit(plan, src, dst)
    plan.stage_timing(True)
    for _ in range(5): plan.execute(src, dst)
    torch.cuda.synchronize()
    st = plan.stage_times()
    plan.stage_timing(False)
    print("   stages(us/exec):", {k: round(v["ms"] * 1e3 / 5, 1) for k, v in st.items() if v["launches"]})
    gs = n / ms / 1e6
    print(f"{name:28s} {ms*1e3:9.1f} us  {gs:8.1f} Gsamples/s  {8*n/ms/1e6:8.1f} GB/s algorithmic  launches={plan.num_launches} ws={plan.workspace_bytes/1e6:.0f}MB", flush=True)
    print(plan.describe())

G3 = gaussian_weights(5.0, 3)
W2 = [[1.0, 0.5, 0.25], [1.0, 0.5, 0.125], [1.0, 0.5, 0.0625], [1.0, 0.5, 0.125], [1.0, 0.5, 0.25], [1.0, 0.5, 0.0625]]
g4 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
run("C1 sat u32 2048^2", (2048, 2048), "u32", [(0, True, [1, 1]), (1, True, [1, 1])])
run("C2 sat f32 4096^2", (4096, 4096), "f32", [(0, True, [1, 1]), (1, True, [1, 1])])
run("C3 gauss 8192^2 fused", (8192, 8192), "f32", g4, "clamp", fuse_dims=1)
run("C3 gauss 8192^2 cascaded", (8192, 8192), "f32", g4, "clamp", fuse_dims=0)
run("C4 audio 64x2^24 r8", (1 << 24, 64), "f32", [(0, True, [1.0] + [0.01] * 8)])
run("C5 512^3", (512, 512, 512), "f32",
    [(0, True, W2[0]), (0, False, W2[1]), (1, True, W2[2]), (1, False, W2[3]), (2, True, W2[4]), (2, False, W2[5])])
