"""One process, several GPUs (rf_mgpu_*, what RecFilter::realize uses with RECFILTER_GPUS=n): strips with a
peer-to-peer tail exchange, or independent parts when the outermost dimension carries no scans.  Needs two
devices (skipped on a one-GPU box; run with `gpurun --gpus 2`).  Results must equal the single-GPU plan to fp32
rounding on every repetition (the exchange is asynchronous: a race would show as run-to-run differences)."""
import math

import numpy as np
import pytest

import recfilter_b200 as rf
from recfilter_b200 import MultiGpuPlan, Plan, Scan, gaussian_weights
from helpers import rand_image, rel_err

pytestmark = pytest.mark.gpu

G3 = gaussian_weights(5.0, 3)
A_ = 2.0 - math.sqrt(3.0)
BIC = [1 + A_, -A_]


def need(n):
    if rf.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("name,ext,dtype,scans,border", [
    ("gaussian", (2048, 2048), "f32", [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)], "clamp"),
    ("bicubic ones", (1024, 1024), "f32", [(0, True, BIC), (0, False, BIC), (1, True, BIC), (1, False, BIC)], "clamp"),
    ("sat u32", (1024, 2048), "u32", [(0, True, [1, 1]), (1, True, [1, 1])], "zero"),
    ("volume z slabs", (128, 128, 256), "f32", [(0, True, [1, .5, .25]), (1, False, [1, .5, .125]), (2, True, [1, .5, .0625]), (2, False, [1, .5, .125])], "zero"),
    ("audio channels (no exchange)", (1 << 16, 8), "f32", [(0, True, [1.0] + [0.01] * 8)], "zero"),
])
def test_two_gpus_equal_one_gpu(oracle, name, ext, dtype, scans, border):
    need(2)
    npdt = np.float32 if dtype == "f32" else np.uint32
    a = np.ones(ext[::-1], npdt) if "ones" in name else rand_image(ext[::-1], npdt, 5)
    if dtype == "u32":
        a = (a % 256).astype(np.uint32)
    one = Plan(ext, dtype, [Scan(*s) for s in scans], border)
    ref = one.realize(a)
    one.close()
    m = MultiGpuPlan(ext, dtype, [Scan(*s) for s in scans], border, ngpus=2)
    assert "2 GPUs" in m.describe()
    for _ in range(6):
        out = m.realize(a)
        if dtype == "u32":
            np.testing.assert_array_equal(out, ref)
        else:
            assert np.isfinite(out).all()
            assert rel_err(out, ref) <= 1e-5, (name, rel_err(out, ref))
    ms = m.profile_host(a, 3)
    assert ms > 0
    m.close()
    if dtype == "f32":
        truth = oracle.apply_filter(a.astype(np.float64), scans, border, threads=8)
        assert rel_err(out, truth) <= 1e-5


def test_plan_is_bound_to_its_device():
    """rf_plan_execute on another device than the plan's fails loudly (the kernel attributes, the workspace and the
    tables belong to the device that was current at rf_plan_create)."""
    need(2)
    import torch
    from recfilter_b200 import RecFilterError
    rf.lib().rf_set_device(0)
    plan = Plan((256, 256), "f32", [Scan(0, True, G3), Scan(1, False, G3)], "clamp")
    a = torch.rand(256 * 256, device="cuda:0")
    plan.execute(a)
    rf.lib().rf_set_device(1)
    try:
        with pytest.raises(RecFilterError, match="created on device 0"):
            plan.execute_ptr(a.data_ptr(), a.data_ptr())
    finally:
        rf.lib().rf_set_device(0)
    plan.close()
