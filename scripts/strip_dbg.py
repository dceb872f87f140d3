import os, sys, math
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from test_fused_gpu import run_sharded, run
from helpers import rel_err
from oracle import oracle
a_ = 2.0 - math.sqrt(3.0)
bic = [1 + a_, -a_]
scans = [(0, True, bic), (0, False, bic), (1, True, bic), (1, False, bic)]
rng = np.random.default_rng(3)
for shape in ((1024, 1024), (2048, 2560)):
    img = rng.random(shape, dtype=np.float32)
    truth = oracle.apply_filter(img.astype(np.float64), scans, "clamp", threads=8)
    for env in ({}, {"RFB_NO_LOCAL_CARRY": "1"}):
        os.environ.update(env)
        whole = run(img, scans, "clamp", engine="twopass")
        sh = run_sharded(img, scans, "clamp", 2, 1, "twopass")
        for k in env: os.environ.pop(k)
        print(shape, env, "whole %.3e" % rel_err(whole, truth), "sharded %.3e" % rel_err(sh, truth), "nan:", np.isnan(sh).sum(), np.isnan(whole).sum(), flush=True)
