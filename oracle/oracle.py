"""Loader for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under recfilter_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.uint32): 3,
       np.dtype(np.int16): 4, np.dtype(np.uint16): 5, np.dtype(np.int8): 6, np.dtype(np.uint8): 7}


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc, separate mul/add, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC",
                               "-shared", "-Wall", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.c_int, C.c_int]
        L.oracle_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_float), C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def apply_filter(array: np.ndarray, scans, border: str = "zero", threads: int = 1, inplace: bool = False):
    """Apply scans (iterable of (dim, causal, coeff)) in order.

    ``array`` has numpy shape ``extents[::-1]`` (dimension 0 = last numpy axis = contiguous),
    ``dim`` counts in the reference's order (0 = x = contiguous).
    """
    a = array if inplace else np.array(array, copy=True, order="C")
    assert a.flags.c_contiguous
    if a.dtype not in _DT:
        raise TypeError(f"oracle: unsupported dtype {a.dtype}")
    scans = list(scans)
    ndim = a.ndim
    ext = (C.c_int64 * ndim)(*[int(e) for e in a.shape[::-1]])
    n = len(scans)
    dims = (C.c_int * max(n, 1))(*[int(s[0]) for s in scans])
    caus = (C.c_int * max(n, 1))(*[1 if s[1] else 0 for s in scans])
    ncs = (C.c_int * max(n, 1))(*[len(s[2]) for s in scans])
    flat = [float(c) for s in scans for c in s[2]]
    coeffs = (C.c_float * max(len(flat), 1))(*flat)
    rc = lib().oracle_filter(a.ctypes.data_as(C.c_void_p), _DT[a.dtype], ndim, ext, 1 if border == "clamp" else 0,
                             n, dims, caus, ncs, coeffs, int(threads))
    if rc != 0:
        raise ValueError("oracle_filter: bad arguments")
    return a
