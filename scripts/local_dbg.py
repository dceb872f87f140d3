import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from recfilter_b200 import Plan, Scan, gaussian_weights
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import rel_err
from oracle import oracle
G3 = gaussian_weights(5.0, 3)
C3 = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
rng = np.random.default_rng(1)
for scans, name in ((C3, "C3"), ([(1, True, G3), (1, False, G3)], "d only"), ([(0, True, G3), (0, False, G3)], "x only"), ([(0, True, G3), (1, True, G3)], "causal xy")):
    a = rng.random((2048, 2560), dtype=np.float32)
    truth = oracle.apply_filter(a.astype(np.float64), scans, "clamp", threads=8)
    res = {}
    for mode in ("local", "chain"):
        if mode == "chain": os.environ["RFB_NO_LOCAL_CARRY"] = "1"
        p = Plan((2560, 2048), "f32", [Scan(*s) for s in scans], "clamp", engine="twopass")
        os.environ.pop("RFB_NO_LOCAL_CARRY", None)
        res[mode] = p.realize(a); d = p.describe().splitlines()[1][-60:]; p.close()
        print(name, mode, "vs truth %.3e" % rel_err(res[mode], truth), d)
    diff = np.abs(res["local"].astype(np.float64) - res["chain"])
    print(name, "local vs chain %.3e" % rel_err(res["local"], res["chain"]), "argmax", np.unravel_index(diff.argmax(), diff.shape))
