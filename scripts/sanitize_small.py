"""Small shapes of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from recfilter_b200 import Plan, Scan, gaussian_weights
G3 = gaussian_weights(5.0, 3)
rng = np.random.default_rng(1)
def run(name, ext, dt, scans, border="zero", **kw):
    a = (rng.random(ext[::-1], dtype=np.float32) if dt == "f32" else rng.integers(0, 255, size=ext[::-1], dtype=np.uint32))
    p = Plan(ext, dt, [Scan(*s) for s in scans], border, **kw)
    out = p.realize(a)
    print(name, "ok", p.describe().splitlines()[1].strip()[:70], float(np.asarray(out, dtype=np.float64).sum()) != 0.0, flush=True)
    p.close()
c3 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
sat = [(0, True, [1, 1]), (1, True, [1, 1])]
run("two-sweep fused 256x384 clamp", (256, 384), "f32", c3, "clamp", engine="twopass")
run("two-sweep fused 64-tile", (128, 192), "f32", c3, "clamp", engine="twopass")
run("ragged two-sweep 200x136", (200, 136), "f32", c3, "clamp", engine="twopass")
run("look-back 2-D u32 256x256", (256, 256), "u32", sat)
run("look-back 2-D f32 order 3", (384, 256), "f32", [(0, False, G3), (1, True, G3)], "clamp")
run("look-back signal r8", (32768, 3), "f32", [(0, True, [1.0] + [0.01] * 8)])
run("look-back signal r3 anticausal", (16384 * 3, 2), "f32", [(0, False, G3)], "clamp")
run("two-sweep signal r8", (32768, 2), "f32", [(0, True, [1.0] + [0.01] * 8)], engine="twopass")
run("generic engine order 5", (96, 80), "f32", [(0, True, [1.0] + [0.1] * 5), (1, False, [1.0] + [0.1] * 5)], engine="generic")
run("3-D volume", (128, 128, 64), "f32", [(0, True, [1, .5, .25]), (1, False, [1, .5, .125]), (2, True, [1, .5, .0625]), (2, False, [1, .5, .125])])
run("fused pointwise epilogue (usm)", (256, 384), "f32", c3, "clamp", epilogue=(2.0, -1.0))
run("fused pointwise epilogue, 128 tiles", (1280, 1024, 4), "f32", c3, "clamp", engine="twopass", epilogue=(2.0, -1.0))
run("order 12, one long line (warp chain, 3 levels)", (65536 + 64,), "f32", [(0, True, [1.0] + [0.01] * 12)])
run("order 20, lines x 40000 (warp chain)", (40000, 3), "f32", [(0, True, [1.0] + [0.01] * 20), (0, False, [1.0] + [0.01] * 20)], "clamp")
os.environ["RFB_STREAM"] = "1"
run("both sweeps in one launch (RFB_STREAM=1)", (1280, 1024, 4), "f32", c3, "clamp", engine="twopass")
del os.environ["RFB_STREAM"]
from recfilter_b200.capi import stencil
import torch
img = torch.rand((300, 520), device="cuda")
box = stencil(img, [(1.0, (3, 3)), (-1.0, (3, -4)), (1.0, (-4, -4)), (-1.0, (-4, 3))], post_scale=1.0 / 49.0)
torch.cuda.synchronize()
print("stencil ok", float(box.sum()) != 0.0, flush=True)
print("SANITIZE SCRIPT DONE")
