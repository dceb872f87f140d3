"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/recfilter_b200.h declares, validates descriptors like the reference's front end
(lib/recfilter.cpp:268-300) and refuses to run without a device (no CPU fallback)."""
import ctypes
import os

import numpy as np
import pytest

import recfilter_b200 as rf
from recfilter_b200 import Plan, Scan, RecFilterError


def test_library_present_and_exports_header_symbols():
    assert os.path.exists(rf.lib_path()), "build the CUDA library first (__graft_entry__.build())"
    L = ctypes.CDLL(rf.lib_path())
    names = rf.exported_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/recfilter_b200.h but not exported"


def test_version_string():
    assert b"sm_100a" in rf.lib().rf_version()


def test_no_cpu_fallback_without_device():
    if rf.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RecFilterError, match="no CUDA device"):
        Plan((64, 64), "f32", [Scan(0, True, [1.0, 0.5])])


@pytest.mark.parametrize("bad", [
    dict(extents=(8, 8), scans=[Scan(2, True, [1.0, 0.5])]),         # unknown dimension
    dict(extents=(8, 8), scans=[Scan(0, True, [1.0])]),               # no feedback coefficient
    dict(extents=(8, 8), scans=[Scan(0, True, [1.0] * 40)]),          # order too high
    dict(extents=(8, 8, 8, 8, 8), scans=[]),                          # too many dimensions
])
def test_descriptor_validation(bad):
    with pytest.raises(RecFilterError):
        Plan(bad["extents"], "f32", bad["scans"])


def test_descriptor_struct_matches_header():
    # sizeof(rf_desc) in the header: 4 + pad + 4*8 + 3*4 + 64*(3*4+33*4) + options(17*4) -> computed by ctypes
    from recfilter_b200.capi import _Desc, _Scan, _Options
    assert ctypes.sizeof(_Scan) == 3 * 4 + 33 * 4
    assert ctypes.sizeof(_Options) == (4 + 5 + 8) * 4
    assert ctypes.sizeof(_Desc) == 8 + 32 + 12 + 64 * ctypes.sizeof(_Scan) + ctypes.sizeof(_Options)


def test_tap_struct_matches_header():
    # rf_tap in include/recfilter_b200.h: float weight; int32 source; int32 offset[4], lo[4], hi[4]
    from recfilter_b200.capi import _Tap
    assert ctypes.sizeof(_Tap) == 4 + 4 + 3 * 4 * 4
    assert _Tap.offset.offset == 8 and _Tap.lo.offset == 24 and _Tap.hi.offset == 40
