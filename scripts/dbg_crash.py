import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from recfilter_b200 import Plan, Scan
W2 = [[1.0, 0.5, 0.25], [1.0, 0.5, 0.125], [1.0, 0.5, 0.0625], [1.0, 0.5, 0.125], [1.0, 0.5, 0.25], [1.0, 0.5, 0.0625]]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "xyz"
sc = []
if "x" in which: sc += [Scan(0, True, W2[0]), Scan(0, False, W2[1])]
if "y" in which: sc += [Scan(1, True, W2[2]), Scan(1, False, W2[3])]
if "z" in which: sc += [Scan(2, True, W2[4]), Scan(2, False, W2[5])]
a = np.random.default_rng(0).random((n, n, n), dtype=np.float32)
p = Plan((n, n, n), "f32", sc, "zero", engine="fused")
print(p.describe())
out = p.realize(a)
print("ok", float(out.mean()))
