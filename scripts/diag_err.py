import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from recfilter_b200 import Plan, Scan, gaussian_weights
from oracle import oracle
from helpers import rand_image, rel_err
G3 = gaussian_weights(5.0, 3)
def run(a, sc, border, **kw):
    p = Plan(a.shape[::-1], a.dtype, [Scan(*s) for s in sc], border, **kw); o = p.realize(a); p.close(); return o
def report(name, a, sc, border, **kw):
    out = run(a, sc, border, **kw)
    truth = oracle.apply_filter(a.astype(np.float64), sc, border)
    ref32 = oracle.apply_filter(a, sc, border)
    print(f"{name:40s} gpu {rel_err(out, truth):.3e}  cpu32 {rel_err(ref32, truth):.3e}  gpu-vs-cpu32 {rel_err(out, ref32):.3e}", flush=True)
for shape in [(64, 64), (136, 200), (1024, 1024)]:
    a = rand_image(shape, np.float32, 30)
    for border in ("zero", "clamp"):
        report(f"{shape} +x {border}", a, [(0, True, G3)], border)
        report(f"{shape} +x-x {border}", a, [(0, True, G3), (0, False, G3)], border)
        report(f"{shape} +y {border}", a, [(1, True, G3)], border)
        report(f"{shape} +y-y {border}", a, [(1, True, G3), (1, False, G3)], border)
        report(f"{shape} 4 scans fused {border}", a, [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)], border, fuse_dims=1)
        report(f"{shape} 4 scans casc {border}", a, [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)], border, fuse_dims=0)
