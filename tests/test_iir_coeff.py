"""Coefficient design against the reference's own lib/iir_coeff.cpp.

tests/golden/iir_coeff_reference.json holds what /root/reference/lib/iir_coeff.cpp -- compiled unchanged by
oracle/Makefile (libiir_ref.so) -- returns over a grid of arguments; oracle/pin_iir.py generated it in the build
container.  Here: recfilter_b200/filters.py (Python) and recfilter_b200/host/iir_coeff.cpp (C++, through
oracle/_ref/libiir_own.so) must reproduce every number to the last bit (float32).
"""
import ctypes
import json
import os

import numpy as np
import pytest

from recfilter_b200 import filters as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "iir_coeff_reference.json")))


def f32(v):
    return [float(np.float32(x)) for x in v]


def test_pin_report():
    rep = json.load(open(os.path.join(ROOT, "tests", "golden", "PIN_IIR_REPORT.json")))
    assert rep["pass"] and rep["max_rel_diff_host_iir_coeff_cpp"] == 0.0 and rep["max_rel_diff_filters_py"] == 0.0


def test_filters_py_equals_reference():
    for r in GOLD["gaussian_weights"]:
        assert f32(F.gaussian_weights(r["sigma"], r["order"])) == r["value"], r
    for r in GOLD["integral_image_coeff"]:
        assert f32(F.integral_image_coeff(r["n"])) == r["value"], r
    for r in GOLD["overlap_feedback_coeff"]:
        assert f32(F.overlap_feedback_coeff(r["a"], r["b"])) == r["value"], r
    for r in GOLD["gaussian_box_filter"]:
        assert int(F.gaussian_box_filter(r["k"], r["sigma"])) == r["value"], r


def test_host_iir_coeff_cpp_equals_reference():
    path = os.path.join(ROOT, "oracle", "_ref", "libiir_own.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libiir_own.so not built (python __graft_entry__.py builds it where /root/reference exists)")
    lib = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.own_gaussian_weights.argtypes = [ctypes.c_float, ctypes.c_int, fp]
    lib.own_integral_image_coeff.argtypes = [ctypes.c_int, fp]
    lib.own_overlap_feedback_coeff.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, fp]
    lib.own_gaussian_box_filter.argtypes = [ctypes.c_int, ctypes.c_float]
    out = (ctypes.c_float * 64)()
    for r in GOLD["gaussian_weights"]:
        n = lib.own_gaussian_weights(r["sigma"], r["order"], out)
        assert f32(out[:n]) == r["value"], r
    for r in GOLD["integral_image_coeff"]:
        n = lib.own_integral_image_coeff(r["n"], out)
        assert f32(out[:n]) == r["value"], r
    for r in GOLD["overlap_feedback_coeff"]:
        a = (ctypes.c_float * len(r["a"]))(*r["a"])
        b = (ctypes.c_float * len(r["b"]))(*r["b"])
        n = lib.own_overlap_feedback_coeff(a, len(r["a"]), b, len(r["b"]), out)
        assert f32(out[:n]) == r["value"], r
    for r in GOLD["gaussian_box_filter"]:
        assert lib.own_gaussian_box_filter(r["k"], r["sigma"]) == r["value"], r
    for name in ("gaussian", "gaussDerivative", "gaussIntegral"):
        fn = getattr(lib, "own_" + name)
        fn.argtypes = [ctypes.c_float] * 3
        fn.restype = ctypes.c_float
        for r in GOLD["point_functions"]:
            assert float(fn(*r["x_mu_sigma"])) == r[name], (name, r)
