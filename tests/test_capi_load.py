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


def test_ctypes_layouts_equal_what_the_c_compiler_sees(tmp_path):
    """sizeof / offsetof of the descriptor structs as gcc lays them out from include/recfilter_b200.h, against the
    ctypes mirrors in recfilter_b200/capi.py (a silent mismatch would shift every option a plan is created with)."""
    import os, shutil, subprocess
    from recfilter_b200.capi import _Desc, _Scan, _Options, _Tap
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "recfilter_b200.h"
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void)
{
    printf("rf_scan %zu\nrf_options %zu\nrf_desc %zu\nrf_tap %zu\n", sizeof(rf_scan), sizeof(rf_options), sizeof(rf_desc), sizeof(rf_tap));
    O(rf_options, honor_tile); O(rf_options, fuse_dims); O(rf_options, open_lo); O(rf_options, open_hi); O(rf_options, shard_dim);
    O(rf_options, engine); O(rf_options, epilogue); O(rf_options, epi_in); O(rf_options, epi_out); O(rf_options, reserved);
    O(rf_desc, extent); O(rf_desc, dtype); O(rf_desc, border); O(rf_desc, nscans); O(rf_desc, scans); O(rf_desc, opt);
    O(rf_scan, causal); O(rf_scan, order); O(rf_scan, coeff);
    O(rf_tap, source); O(rf_tap, offset); O(rf_tap, lo); O(rf_tap, hi);
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirror = {"rf_scan": _Scan, "rf_options": _Options, "rf_desc": _Desc, "rf_tap": _Tap}
    for key, val in got.items():
        if "." in key:
            st, field = key.split(".")
            assert getattr(mirror[st], field).offset == int(val), key
        else:
            assert ctypes.sizeof(mirror[key]) == int(val), key
