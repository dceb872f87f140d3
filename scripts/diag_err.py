"""Development aid: error of the fused / generic engines and of the serial fp32 loop against the
fp64 truth (max |d| / max |truth|), for the cases whose precision the carry algebra decides."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from recfilter_b200 import Plan, Scan, gaussian_weights
from oracle import oracle
from helpers import rand_image, rel_err
G3 = gaussian_weights(5.0, 3)
G2 = gaussian_weights(5.0, 2)
def run(a, sc, border, **kw):
    p = Plan(a.shape[::-1], a.dtype, [Scan(*s) for s in sc], border, **kw); o = p.realize(a); p.close(); return o
def report(name, a, sc, border):
    truth = oracle.apply_filter(a.astype(np.float64), sc, border, threads=8)
    ref32 = oracle.apply_filter(a, sc, border, threads=8)
    f = run(a, sc, border, engine="fused")
    g = run(a, sc, border, engine="generic")
    print(f"{name:34s} fused {rel_err(f, truth):.2e}  generic {rel_err(g, truth):.2e}  cpu32 {rel_err(ref32, truth):.2e}  "
          f"fused-generic {rel_err(f, g):.2e}", flush=True)
C3 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
SAT = [(0, True, [1, 1]), (1, True, [1, 1])]
SAT2 = [(0, True, [1, 2, -1]), (1, True, [1, 2, -1])]
for shape in [(128, 128), (256, 384), (1024, 1024), (2048, 2048), (4096, 4096)]:
    a = rand_image(shape, np.float32, 300)
    for border in ("zero", "clamp"):
        report(f"{shape} C3 {border}", a, C3, border)
    report(f"{shape} SAT f32", a, SAT, "zero")
    report(f"{shape} SAT r=2 f32", a, SAT2, "zero")
    report(f"{shape} G2 xy", a, [(0, True, G2), (0, False, G2), (1, True, G2), (1, False, G2)], "clamp")
