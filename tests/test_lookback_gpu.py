"""GPU parity tests of the single-pass decoupled look-back kernels (recfilter_b200/csrc/lookback.cuh)
through the C ABI: lb_tile_kernel (2-D, at most one scan per dimension) and lb_signal_kernel (long 1-D
signals).  The planner must pick them for one-way filters ("single-pass" in the plan text), results are
compared with the oracle (bit exact for integers, 1e-5 of max|truth| for fp32 against the fp64 oracle) and
with the two-sweep kernels (engine="twopass").  Reference semantics: lib/recfilter.cpp:260-392; the serial
inter-tile loop these kernels replace: lib/split.cpp:832-846.
"""
import os

import numpy as np
import pytest

from recfilter_b200 import Plan, Scan, gaussian_weights
from helpers import rand_image, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5
SAT = [(0, True, [1, 1]), (1, True, [1, 1])]
G3 = gaussian_weights(5.0, 3)
G2 = gaussian_weights(5.0, 2)
A8 = [1.0] + [0.01] * 8                                   # apps/audio/audio_filter_high_order.cpp:41-42
B8 = [0.2, 0.9, -0.5, 0.3, -0.2, 0.1, -0.05, 0.02, -0.01]   # stable order-8 set with a non-unit feed-forward


def run(a, scans, border="zero", engine="auto", expect="single-pass", ts=None, **kw):
    if ts is not None:
        os.environ["RFB_LB_TS"] = str(ts)
    try:
        plan = Plan(a.shape[::-1], a.dtype, [Scan(*s) for s in scans], border, engine=engine, **kw)
    finally:
        os.environ.pop("RFB_LB_TS", None)
    if expect:
        assert expect in plan.describe(), plan.describe()
    out = plan.realize(a)            # rf_plan_execute_host ends with rf_plan_check: a spin-limit failure raises
    plan.close()
    return out


def check_float(oracle, a, scans, border="zero", tol=TOL, **kw):
    out = run(a, scans, border, **kw)
    truth = oracle.apply_filter(a.astype(np.float64), scans, border, threads=8)
    ref32 = oracle.apply_filter(a, scans, border, threads=8)
    e_gpu, e_cpu = rel_err(out, truth), rel_err(ref32, truth)
    assert np.isfinite(out).all()
    assert e_gpu <= tol, f"look-back rel err {e_gpu:.3e} (serial fp32 loop {e_cpu:.3e})"
    assert e_gpu <= 8 * e_cpu + 2e-6, f"look-back {e_gpu:.3e} much worse than the serial fp32 loop {e_cpu:.3e}"
    return out


# ---- C1: summed-area table, uint32, bit exact ----
@pytest.mark.parametrize("n", [64, 128, 512, 2048])
@pytest.mark.parametrize("ts", [64, 128])
def test_c1_sat_u32_bit_exact(oracle, n, ts):
    if n % ts:
        pytest.skip("extent not a multiple of the tile")
    rng = np.random.default_rng(20240601)
    a = rng.integers(0, 256, size=(n, n), dtype=np.uint32)
    out = run(a, SAT, ts=ts)
    np.testing.assert_array_equal(out, oracle.apply_filter(a, SAT, threads=8))
    if n == 2048:
        ones = run(np.ones((n, n), np.uint32), SAT, ts=ts)
        yy, xx = np.mgrid[0:n, 0:n]
        np.testing.assert_array_equal(ones, ((xx + 1) * (yy + 1)).astype(np.uint32))
        full = rand_image((n, n), np.uint32, 21)            # full-range: wraparound must match
        np.testing.assert_array_equal(run(full, SAT, ts=ts), oracle.apply_filter(full, SAT, threads=8))


def test_c1_repeated_launches_are_identical():
    # the look-back walk depends on timing; integer results must not
    import torch
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << 32, size=(2048, 2048), dtype=np.uint32)
    plan = Plan((2048, 2048), "u32", [Scan(*s) for s in SAT])
    assert "single-pass" in plan.describe()
    src = torch.from_numpy(a.view(np.int32)).cuda()
    first = plan.execute(src).clone()
    for _ in range(30):
        assert torch.equal(plan.execute(src), first)
    plan.check()
    ref = np.cumsum(np.cumsum(a.astype(np.uint64), axis=0) & 0xFFFFFFFF, axis=1) & 0xFFFFFFFF
    np.testing.assert_array_equal(first.cpu().numpy().view(np.uint32), ref.astype(np.uint32))
    plan.close()


# ---- C2: the box filters' integral images (apps/box/box_filter.h:29-39, 118-139), fp32 ----
def test_c2_sat_f32_full_size(oracle):
    a = rand_image((4096, 4096), np.float32, 1)
    out = check_float(oracle, a, SAT)
    assert out[-1, -1] == pytest.approx(float(a.astype(np.float64).sum()), rel=1e-5)
    two = run(a, SAT, engine="twopass", expect="fused pass")
    assert rel_err(out, two) < 2 * TOL


@pytest.mark.parametrize("ts", [64, 128])
def test_second_order_integral(oracle, ts):
    b = rand_image((256, 384), np.float32, 41)
    # double pole at 1: the serial fp32 loop itself sits at 3.5e-5 here
    check_float(oracle, b, [(0, True, [1, 2, -1]), (1, True, [1, 2, -1])], tol=4e-5, ts=ts)


# ---- every direction, orders 1..4, borders, single-dimension filters ----
@pytest.mark.parametrize("cx,cd", [(True, True), (False, True), (True, False), (False, False)])
@pytest.mark.parametrize("border", ["zero", "clamp"])
@pytest.mark.parametrize("ts", [64, 128])
def test_directions_and_borders(oracle, cx, cd, border, ts):
    a = rand_image((384, 512), np.float32, 7) - np.float32(0.5)
    check_float(oracle, a, [(0, cx, G3), (1, cd, G2)], border, ts=ts)


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_orders(oracle, order):
    coeff = {1: gaussian_weights(5.0, 1), 2: G2, 3: G3, 4: [0.3, 0.8, -0.3, 0.15, -0.05]}[order]
    a = rand_image((256, 640), np.float32, 50 + order)
    check_float(oracle, a, [(0, True, coeff), (1, False, coeff)], "clamp")


@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("causal", [True, False])
def test_single_dimension(oracle, dim, causal):
    a = rand_image((512, 384), np.float32, 60)
    check_float(oracle, a, [(dim, causal, G3)], "clamp")
    u = rand_image((256, 256), np.uint32, 61)
    np.testing.assert_array_equal(run(u, [(dim, causal, [1, 3, -1])]), oracle.apply_filter(u, [(dim, causal, [1, 3, -1])]))


def test_many_tiles_per_line(oracle):
    # long look-back walks: 128 tiles along x, 2 along d (and the transpose)
    a = rand_image((128, 8192), np.float32, 62)
    check_float(oracle, a, [(0, True, G3), (1, True, G3)], "clamp", ts=64)
    check_float(oracle, a.T.copy(), [(0, False, G3), (1, False, G3)], "zero", ts=64)
    u = rand_image((128, 8192), np.uint32, 63)
    np.testing.assert_array_equal(run(u, SAT, ts=64), oracle.apply_filter(u, SAT, threads=8))


def test_int_types(oracle):
    for dt in (np.int32, np.uint16, np.int8):
        a = rand_image((128, 192 + 64), dt, 90)
        sc = [(0, False, [1, -1, 3]), (1, True, [1, 2, -1])]
        np.testing.assert_array_equal(run(a, sc), oracle.apply_filter(a, sc))


def test_stack_of_images(oracle):
    # outermost dimension without scans (lib/split.cpp:1888-1898): every image of the stack is filtered alone
    stack = rand_image((3, 256, 384), np.float32, 4242)
    sc = [(0, True, G2), (1, True, G2)]
    out = run(stack, sc, "clamp")
    for b in range(3):
        truth = oracle.apply_filter(stack[b].astype(np.float64), sc, "clamp")
        assert rel_err(out[b], truth) <= TOL


def test_two_scans_per_dimension_keep_the_two_sweep_kernels():
    plan = Plan((256, 256), "f32", [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3)], "clamp")
    assert "single-pass" not in plan.describe()
    plan.close()


# ---- long 1-D signals ----
@pytest.mark.parametrize("n,rows", [(16384, 3), (131072, 4), (1 << 20, 1), (1 << 18, 5)])
@pytest.mark.parametrize("coeff", [A8, B8, G3, [1.0, 1.0]], ids=["ref8", "stable8", "gauss3", "sum"])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("tile_rows", [32, 64, 128])
def test_signal_lookback_matches_oracle(oracle, n, rows, coeff, causal, tile_rows):
    a = rand_image((rows, n), np.float32, 77) - np.float32(0.5)
    scans = [(0, causal, coeff)]
    for border in ("zero", "clamp"):
        os.environ["RFB_LB_ROWS"] = str(tile_rows)          # rows of 128 samples per CTA (1, 2 or 4 warps)
        try:
            plan = Plan((n, rows), "f32", [Scan(*s) for s in scans], border)
        finally:
            os.environ.pop("RFB_LB_ROWS", None)
        assert f"{tile_rows} rows per CTA" in plan.describe() or "look-back signal pass" not in plan.describe()
        if len(coeff) - 1 > 4 or n // 128 > 128:          # otherwise the 2-D kernel takes it (few tiles per line)
            assert "look-back signal pass" in plan.describe(), plan.describe()
        else:
            assert "single-pass" in plan.describe(), plan.describe()
        out = plan.realize(a)
        plan.close()
        truth = oracle.apply_filter(a.astype(np.float64), scans, border, threads=8)
        assert np.isfinite(out).all()
        assert rel_err(out, truth) <= TOL, f"{border}: {rel_err(out, truth):.3e}"


def test_signal_lookback_u32_prefix_sum_bit_exact(oracle):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 1 << 32, size=(3, 40 * 16384), dtype=np.uint32)          # wraps many times, 40 tiles per signal
    for causal in (True, False):
        scans = [(0, causal, [1, 1])]
        plan = Plan((40 * 16384, 3), "u32", [Scan(*s) for s in scans])
        assert "look-back signal pass" in plan.describe(), plan.describe()
        np.testing.assert_array_equal(plan.realize(a), oracle.apply_filter(a, scans, threads=8))
        plan.close()


def test_c4_audio_full_size(oracle):
    """BASELINE C4 at its full size: 64 channels x 2^24 samples, order 8 (apps/audio/audio_filter_high_order.cpp:41-55),
    checked channel by channel against the fp64 oracle (the whole array in fp64 would be 8.6 GB)."""
    import torch
    n, ch = 1 << 24, 64
    plan = Plan((n, ch), "f32", [Scan(0, True, A8)])
    assert "look-back signal pass" in plan.describe(), plan.describe()
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.rand((ch, n), device="cuda", generator=g) * 2 - 1
    dst = plan.execute(src)
    plan.check()
    two = Plan((n, ch), "f32", [Scan(0, True, A8)], engine="twopass")
    assert "signal pass" in two.describe() and "look-back" not in two.describe()
    dst2 = two.execute(src)
    torch.cuda.synchronize()
    worst = 0.0
    for c in range(0, ch, 7):                               # 10 channels spread over the array, every tile position
        x = src[c].cpu().numpy()
        truth = oracle.apply_filter(x.astype(np.float64)[None, :], [(0, True, A8)], threads=1)[0]
        worst = max(worst, rel_err(dst[c].cpu().numpy(), truth), rel_err(dst2[c].cpu().numpy(), truth))
    assert worst <= TOL, worst
    # the two engines agree everywhere (cheap full-array check on the device)
    scale = float(dst2.abs().max())
    assert float((dst - dst2).abs().max()) / scale <= 2 * TOL
    plan.close(); two.close()


def test_short_memory_specialisations_equal_the_general_path(oracle):
    """Short-memory filters: the tail pass skips the chunks of a row that do not reach its tail, and the carry of a
    tile is the aggregate of the tile before it (both decided from fp64 bounds < 1e-12).  Same result as the general
    path (RFB_NO_SHORT_MEMORY=1) far below the tolerance; a prefix sum (pole 1) must keep the general path."""
    a = rand_image((3, 1 << 18), np.float32, 99) - np.float32(0.5)
    for coeff in (A8, B8, G3):
        scans = [(0, True, coeff)]
        plan = Plan((1 << 18, 3), "f32", [Scan(*s) for s in scans])
        assert "carries from the previous tile only" in plan.describe(), plan.describe()
        fast = plan.realize(a)
        plan.close()
        os.environ["RFB_NO_SHORT_MEMORY"] = "1"
        try:
            plan = Plan((1 << 18, 3), "f32", [Scan(*s) for s in scans])
            assert "previous tile only" not in plan.describe() and "short memory" not in plan.describe()
            general = plan.realize(a)
            plan.close()
        finally:
            os.environ.pop("RFB_NO_SHORT_MEMORY", None)
        assert rel_err(fast, general) <= 2e-7, rel_err(fast, general)
        truth = oracle.apply_filter(a.astype(np.float64), scans, threads=8)
        assert rel_err(fast, truth) <= TOL
    plan = Plan((1 << 18, 3), "f32", [Scan(0, True, [1.0, 1.0])])
    assert "previous tile only" not in plan.describe() and "short memory" not in plan.describe(), plan.describe()
    plan.close()


def test_lookback_plans_replay_in_a_cuda_graph(oracle):
    """Nothing about a launch is baked into the kernel parameters (the epoch of the record tags lives in device
    memory): a captured graph of a look-back plan must give the right answer for NEW inputs on every replay."""
    import torch
    sat = Plan((1024, 512), "u32", [Scan(*s) for s in SAT])
    sig = Plan((1 << 18, 4), "f32", [Scan(0, True, B8)])
    assert "single-pass" in sat.describe() and "look-back signal pass" in sig.describe()
    a = torch.zeros((512, 1024), device="cuda", dtype=torch.int32); ao = torch.empty_like(a)
    b = torch.zeros((4, 1 << 18), device="cuda", dtype=torch.float32); bo = torch.empty_like(b)
    sat.execute(a, ao); sig.execute(b, bo)                    # lazy initialisations outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sat.execute(a, ao)
        sig.execute(b, bo)
    rng = np.random.default_rng(12)
    for rep in range(4):
        an = rng.integers(0, 1 << 31, size=(512, 1024), dtype=np.int64).astype(np.uint32)
        bn = rng.random((4, 1 << 18), dtype=np.float32) - np.float32(0.5)
        a.copy_(torch.from_numpy(an.view(np.int32))); b.copy_(torch.from_numpy(bn))
        g.replay()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(ao.cpu().numpy().view(np.uint32), oracle.apply_filter(an, SAT, threads=8))
        truth = oracle.apply_filter(bn.astype(np.float64), [(0, True, B8)], threads=8)
        assert rel_err(bo.cpu().numpy(), truth) <= TOL
    sat.check(); sig.check()
    sat.close(); sig.close()


def test_several_scans_along_long_lines_run_one_signal_pass_each(oracle):
    """apps/audio/audio_filter_biquads.cpp: up to a dozen causal order-2 scans on one signal.  No tile engine fuses them,
    so the planner emits one single-pass signal kernel per scan (the passes run in place, one after the other)."""
    biquad = [1.0, 0.1, 0.1]
    a = (rand_image((2, 1 << 17), np.float32, 91) * 2 - 1).astype(np.float32)
    scans = [(0, True, biquad)] * 6
    plan = Plan(a.shape[::-1], a.dtype, [Scan(*s) for s in scans])
    text = plan.describe()
    assert text.count("look-back signal pass") == 6 and plan.num_launches == 6, text
    out = plan.realize(a)
    plan.close()
    truth = oracle.apply_filter(a.astype(np.float64), scans, threads=8)
    assert rel_err(out, truth) <= 1e-5
    # causal and anticausal scans of different orders, clamped border
    mixed = [(0, True, [0.3, 0.4, 0.2, 0.05]), (0, False, [0.5, 0.3, 0.1]), (0, True, biquad), (0, False, [0.6, 0.35]), (0, True, biquad)]
    plan = Plan(a.shape[::-1], a.dtype, [Scan(*s) for s in mixed], "clamp")
    assert plan.describe().count("look-back signal pass") == 5
    out = plan.realize(a)
    plan.close()
    truth = oracle.apply_filter(a.astype(np.float64), mixed, "clamp", threads=8)
    assert rel_err(out, truth) <= 1e-5
    # integer ring: bit exact
    b = rand_image((3, 1 << 16), np.uint32, 92)
    isc = [(0, True, [1, 1]), (0, True, [1, 2, -1]), (0, False, [1, 1]), (0, True, [1, 3]), (0, False, [1, -1, 2])]
    plan = Plan(b.shape[::-1], b.dtype, [Scan(*s) for s in isc])
    assert plan.describe().count("look-back signal pass") == 5
    assert np.array_equal(plan.realize(b), oracle.apply_filter(b, isc))
    plan.close()
