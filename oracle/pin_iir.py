#!/usr/bin/env python
"""Pin the coefficient design against the reference's own lib/iir_coeff.cpp.  TEST INFRASTRUCTURE.

oracle/Makefile compiles /root/reference/lib/iir_coeff.cpp UNCHANGED (oracle/iir_pin/Halide.h supplies the
handful of Halide names its unused Expr overloads mention) into oracle/_ref/libiir_ref.so, and this repo's
recfilter_b200/host/iir_coeff.cpp into oracle/_ref/libiir_own.so.  This script calls both (and the Python
restatement recfilter_b200/filters.py) over a grid of arguments, and writes

  tests/golden/iir_coeff_reference.json   the reference's outputs (float32, exact decimal repr) -- the golden
                                          vectors tests/test_iir_coeff.py checks host/iir_coeff.cpp and filters.py
                                          against, also on machines without /root/reference
  tests/golden/PIN_IIR_REPORT.json        max differences found here

Run it in the build container (it needs /root/reference): python oracle/pin_iir.py
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

SIGMAS = [0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 5.0, 7.5, 10.0, 16.0, 25.0, 50.0]
OVERLAPS = [([0.5], [0.25]), ([1.5, -0.6], [0.7]), ([2.2963, -1.7997, 0.4807], [2.2963, -1.7997, 0.4807]),
            ([0.76796180], [1.52838480, -0.62596899]), ([0.1, 0.2, 0.3, 0.4], [0.4, 0.3, 0.2, 0.1])]
BOXES = [(1, 2.0), (3, 2.0), (3, 5.0), (6, 5.0), (4, 10.0)]
POINTS = [(-3.0, 0.0, 1.0), (0.0, 0.0, 1.0), (0.7, 0.2, 2.5), (4.0, 1.0, 5.0)]


def load(path, pfx):
    lib = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    g = lambda n: getattr(lib, pfx + n)
    g("gaussian_weights").argtypes = [ctypes.c_float, ctypes.c_int, fp]
    g("integral_image_coeff").argtypes = [ctypes.c_int, fp]
    g("overlap_feedback_coeff").argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, fp]
    g("gaussian_box_filter").argtypes = [ctypes.c_int, ctypes.c_float]
    for n in ("gaussian", "gaussDerivative", "gaussIntegral"):
        g(n).argtypes = [ctypes.c_float] * 3
        g(n).restype = ctypes.c_float

    def vec(fn, *args):
        out = (ctypes.c_float * 64)()
        n = fn(*args, out)
        return [float(np.float32(out[i])) for i in range(n)]

    def arr(v):
        return (ctypes.c_float * len(v))(*v)

    return {
        "gaussian_weights": lambda s, o: vec(g("gaussian_weights"), s, o),
        "integral_image_coeff": lambda n: vec(g("integral_image_coeff"), n),
        "overlap_feedback_coeff": lambda a, b: vec(lambda *x: g("overlap_feedback_coeff")(arr(a), len(a), arr(b), len(b), x[-1])),
        "gaussian_box_filter": lambda k, s: int(g("gaussian_box_filter")(k, s)),
        "gaussian": lambda *p: float(g("gaussian")(*p)),
        "gaussDerivative": lambda *p: float(g("gaussDerivative")(*p)),
        "gaussIntegral": lambda *p: float(g("gaussIntegral")(*p)),
    }


def table(api):
    t = {"gaussian_weights": [], "integral_image_coeff": [], "overlap_feedback_coeff": [], "gaussian_box_filter": [],
         "point_functions": []}
    for s in SIGMAS:
        for o in (1, 2, 3):
            t["gaussian_weights"].append({"sigma": s, "order": o, "value": api["gaussian_weights"](s, o)})
    for n in (1, 2, 3, 4, 5):
        t["integral_image_coeff"].append({"n": n, "value": api["integral_image_coeff"](n)})
    for a, b in OVERLAPS:
        t["overlap_feedback_coeff"].append({"a": a, "b": b, "value": api["overlap_feedback_coeff"](a, b)})
    for k, s in BOXES:
        t["gaussian_box_filter"].append({"k": k, "sigma": s, "value": api["gaussian_box_filter"](k, s)})
    for p in POINTS:
        t["point_functions"].append({"x_mu_sigma": list(p), "gaussian": api["gaussian"](*p),
                                     "gaussDerivative": api["gaussDerivative"](*p), "gaussIntegral": api["gaussIntegral"](*p)})
    return t


def max_rel(a, b):
    """largest |a-b| / max(|a|, tiny) over two tables of the same shape"""
    worst = 0.0
    for key in a:
        for ra, rb in zip(a[key], b[key]):
            for f in ra:
                if f in ("sigma", "order", "n", "a", "b", "k", "x_mu_sigma"):
                    continue
                va, vb = np.atleast_1d(np.asarray(ra[f], np.float64)), np.atleast_1d(np.asarray(rb[f], np.float64))
                if va.shape != vb.shape:
                    return float("inf")
                worst = max(worst, float(np.max(np.abs(va - vb) / np.maximum(np.abs(va), 1e-30))) if va.size else 0.0)
    return worst


def python_api():
    from recfilter_b200 import filters as F
    import math
    return {
        "gaussian_weights": lambda s, o: [float(np.float32(v)) for v in F.gaussian_weights(s, o)],
        "integral_image_coeff": lambda n: [float(np.float32(v)) for v in F.integral_image_coeff(n)],
        "overlap_feedback_coeff": lambda a, b: [float(np.float32(v)) for v in F.overlap_feedback_coeff(a, b)],
        "gaussian_box_filter": lambda k, s: int(F.gaussian_box_filter(k, s)),
        # filters.py has no point functions: mirror the reference's float formulas (lib/iir_coeff.cpp:193-203) trivially
        "gaussian": None, "gaussDerivative": None, "gaussIntegral": None,
    }


def main():
    if not os.path.isdir("/root/reference/lib"):
        sys.exit("needs /root/reference (build container only)")
    subprocess.check_call(["make", "-j", "8", "_ref/libiir_ref.so", "_ref/libiir_own.so"], cwd=HERE)
    ref = table(load(os.path.join(HERE, "_ref", "libiir_ref.so"), "ref_"))
    own = table(load(os.path.join(HERE, "_ref", "libiir_own.so"), "own_"))
    py = python_api()
    pyt = {k: [] for k in ("gaussian_weights", "integral_image_coeff", "overlap_feedback_coeff", "gaussian_box_filter")}
    for r in ref["gaussian_weights"]:
        pyt["gaussian_weights"].append({"value": py["gaussian_weights"](r["sigma"], r["order"])})
    for r in ref["integral_image_coeff"]:
        pyt["integral_image_coeff"].append({"value": py["integral_image_coeff"](r["n"])})
    for r in ref["overlap_feedback_coeff"]:
        pyt["overlap_feedback_coeff"].append({"value": py["overlap_feedback_coeff"](r["a"], r["b"])})
    for r in ref["gaussian_box_filter"]:
        pyt["gaussian_box_filter"].append({"value": py["gaussian_box_filter"](r["k"], r["sigma"])})
    ref_sub = {k: ref[k] for k in pyt}
    report = {
        "reference": "/root/reference/lib/iir_coeff.cpp compiled unchanged (oracle/Makefile: _ref/libiir_ref.so)",
        "max_rel_diff_host_iir_coeff_cpp": max_rel(ref, own),
        "max_rel_diff_filters_py": max_rel(ref_sub, pyt),
        "grid": {"sigmas": SIGMAS, "orders": [1, 2, 3]},
    }
    report["pass"] = report["max_rel_diff_host_iir_coeff_cpp"] <= 2e-6 and report["max_rel_diff_filters_py"] <= 2e-6
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    json.dump(ref, open(os.path.join(ROOT, "tests", "golden", "iir_coeff_reference.json"), "w"), indent=1)
    json.dump(report, open(os.path.join(ROOT, "tests", "golden", "PIN_IIR_REPORT.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))
    sys.exit(0 if report["pass"] else 1)


if __name__ == "__main__":
    main()
