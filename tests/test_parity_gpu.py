"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on seeded inputs.

Integer filters must be bit exact.  fp32 filters are compared with the oracle run in
float64 on the same float32 input and float32 coefficients ("truth"); tolerance
1e-5 relative to the output scale (BASELINE.json north_star), and the GPU must not be
further from the truth than 8x the serial fp32 loop is (plus a floor)."""
import numpy as np
import pytest

import recfilter_b200 as rf
from recfilter_b200 import Plan, Scan, gaussian_weights
from helpers import rand_image, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5


def run_gpu(a, scans, border="zero", **kw):
    ext = a.shape[::-1]
    plan = Plan(ext, a.dtype, [Scan(*s) for s in scans], border, **kw)
    out = plan.realize(a)
    plan.close()
    return out


def check_float(oracle, a, scans, border="zero", tol=TOL, **kw):
    out = run_gpu(a, scans, border, **kw)
    truth = oracle.apply_filter(a.astype(np.float64), scans, border)
    ref32 = oracle.apply_filter(a, scans, border)
    e_gpu = rel_err(out, truth)
    e_cpu = rel_err(ref32, truth)
    assert np.isfinite(out).all()
    assert e_gpu <= tol, f"gpu rel err {e_gpu:.3e} (serial fp32 loop: {e_cpu:.3e})"
    assert e_gpu <= 8 * e_cpu + 2e-6, f"gpu {e_gpu:.3e} much worse than serial fp32 {e_cpu:.3e}"
    return out


def check_int(oracle, a, scans, border="zero", **kw):
    out = run_gpu(a, scans, border, **kw)
    ref = oracle.apply_filter(a, scans, border)
    np.testing.assert_array_equal(out, ref)
    return out


W2 = [[1.0, 0.5, 0.25], [1.0, 0.5, 0.125], [1.0, 0.5, 0.0625], [1.0, 0.5, 0.125], [1.0, 0.5, 0.25], [1.0, 0.5, 0.0625]]
G3 = gaussian_weights(5.0, 3)


# ---- the reference's own test configurations, with seeded random input (tile 4 honoured) ----------------
def test_ref_trivial_sat(oracle):                       # tests/test_trivial.cpp
    a = rand_image((20, 20), np.float32, 10)
    check_float(oracle, a, [(0, True, [1, 1]), (1, True, [1, 1])], tile=4, honor_tile=True)


def test_ref_type_invariance_int16(oracle):             # tests/test_type_invariance.cpp
    a = rand_image((20, 20), np.int16, 11)
    check_int(oracle, a, [(0, True, [1, 1, -1]), (1, True, [1, 1, -1])], tile=4, honor_tile=True)


def test_ref_repeated_causal(oracle):                   # tests/test_repeated_causal.cpp: 4x +x order 3
    a = rand_image((1, 20), np.float32, 12)
    W = [[1, .5, .25, .125], [1, .5, .125, .0625], [1, .25, .125, .0625], [1, .125, .0625, .03125]]
    check_float(oracle, a, [(0, True, w) for w in W], tile=4, honor_tile=True)


def test_ref_repeated_anticausal(oracle):               # tests/test_repeated_anticausal.cpp: 4x -x order 2
    a = rand_image((1, 20), np.float32, 13)
    check_float(oracle, a, [(0, False, w) for w in W2[:4]], tile=4, honor_tile=True)


def test_ref_causal_anticausal(oracle):                 # tests/test_causal_anticausal.cpp (+x,-x,+x,-x)
    a = rand_image((1, 20), np.float32, 14)
    w = [0.5, 0.0625, 0.0]
    check_float(oracle, a, [(0, True, [1] + w), (0, False, [1] + w), (0, True, [1] + w), (0, False, [1] + w)],
                tile=4, honor_tile=True)


def test_ref_causal_xy(oracle):                         # tests/test_causal_xy.cpp
    a = rand_image((16, 16), np.float32, 15)
    w = [1, .5, .25, .125]
    check_float(oracle, a, [(0, True, w), (0, True, w), (1, True, w), (1, True, w)], tile=4, honor_tile=True)


def test_ref_causal_anticausal_xy(oracle):              # tests/test_causal_anticausal_xy.cpp
    a = rand_image((16, 16), np.float32, 16)
    w = [1, .5, .25, .125]
    check_float(oracle, a, [(0, True, w), (0, False, w), (1, True, w), (1, False, w)], tile=4, honor_tile=True)


def test_ref_generic_xy(oracle):                        # tests/test_generic_xy.cpp: +x,-x,+x,-x,+y,-y,-y
    a = rand_image((16, 16), np.float32, 17)
    sc = [(0, True, W2[0]), (0, False, W2[1]), (0, True, W2[2]), (0, False, W2[3]),
          (1, True, W2[4]), (1, False, W2[5]), (1, False, W2[0])]
    check_float(oracle, a, sc, tile=4, honor_tile=True)


def test_ref_generic_xyz(oracle):                       # tests/test_generic_xyz.cpp
    a = rand_image((16, 16, 16), np.float32, 18)
    sc = [(0, True, W2[0]), (0, False, W2[1]), (1, True, W2[2]), (1, False, W2[3]), (2, True, W2[4]), (2, False, W2[5])]
    check_float(oracle, a, sc, tile=4, honor_tile=True)


# ---- BASELINE.json configurations at reduced and full size ----------------------------------------------
@pytest.mark.parametrize("n", [64, 200, 2048])
def test_c1_sat_u32_bit_exact(oracle, n):
    rng = np.random.default_rng(20240601)
    a = rng.integers(0, 256, size=(n, n), dtype=np.uint32)
    out = check_int(oracle, a, [(0, True, [1, 1]), (1, True, [1, 1])])
    if n == 2048:
        ones = np.ones((n, n), np.uint32)
        sat = run_gpu(ones, [(0, True, [1, 1]), (1, True, [1, 1])])
        yy, xx = np.mgrid[0:n, 0:n]
        np.testing.assert_array_equal(sat, ((xx + 1) * (yy + 1)).astype(np.uint32))


def test_c1_full_range_wraparound(oracle):
    a = rand_image((300, 333), np.uint32, 21)
    check_int(oracle, a, [(0, True, [1, 1]), (1, True, [1, 1])])


@pytest.mark.parametrize("shape", [(256, 256), (136, 200), (70, 1000), (1, 300), (300, 1), (64, 64), (65, 63)])
@pytest.mark.parametrize("fuse", [1, 0])
def test_c3_gaussian_clamped(oracle, shape, fuse):
    a = rand_image(shape, np.float32, 30)
    sc = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
    check_float(oracle, a, sc, "clamp", fuse_dims=fuse)


def test_c3_gaussian_constant_image_fixed_point(oracle):
    a = np.full((512, 512), 0.625, np.float32)
    sc = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
    out = run_gpu(a, sc, "clamp")
    np.testing.assert_allclose(out, 0.625, rtol=3e-4)


def test_c3_fused_equals_cascaded(oracle):
    a = rand_image((300, 520), np.float32, 31)
    sc = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
    f = run_gpu(a, sc, "clamp", fuse_dims=1)
    c = run_gpu(a, sc, "clamp", fuse_dims=0)
    assert rel_err(f, c) < 8e-6      # two independent fp32 roundings of the same filter


def test_c2_box_core_sat_and_order2(oracle):
    a = rand_image((256, 384), np.float32, 40)
    check_float(oracle, a, [(0, True, [1, 1]), (1, True, [1, 1])])
    b = rand_image((128, 192), np.float32, 41)
    check_float(oracle, b, [(0, True, [1, 2, -1]), (1, True, [1, 2, -1])], tol=2e-5)


def test_c4_audio_order8(oracle):
    a = (rand_image((64, 1 << 14), np.float32, 50) * 2 - 1).astype(np.float32)
    coeff = [1.0] + [0.01] * 8
    check_float(oracle, a, [(0, True, coeff)])


def test_c4_long_signal_two_level_chain(oracle):
    # 1-D signal: signal mode, more than 512 tiles -> segmented carry chain
    a = (rand_image((1, 100003), np.float32, 51) * 2 - 1).astype(np.float32)
    check_float(oracle, a, [(0, True, [1.0] + [0.01] * 8)])
    check_float(oracle, a, [(0, True, [0.2, 0.5, 0.2, 0.05]), (0, False, [0.2, 0.5, 0.2, 0.05])])
    b = rand_image((3, 70001), np.uint32, 52)
    check_int(oracle, b, [(0, True, [1, 1]), (0, False, [1, 2, -1])])


def test_c4_many_rows_long(oracle):
    a = rand_image((70, 40000), np.float32, 53)
    check_float(oracle, a, [(0, True, [0.3, 0.4, 0.2, 0.05]), (0, False, [0.5, 0.3, 0.1])], "clamp")


@pytest.mark.parametrize("shape", [(40, 36, 50), (70, 64, 130), (3, 200, 5)])
def test_c5_volume(oracle, shape):
    a = rand_image(shape, np.float32, 60)
    sc = [(0, True, W2[0]), (0, False, W2[1]), (1, True, W2[2]), (1, False, W2[3]), (2, True, W2[4]), (2, False, W2[5])]
    check_float(oracle, a, sc)
    check_float(oracle, a, sc, "clamp", fuse_dims=0)


def test_4d(oracle):
    a = rand_image((5, 6, 70, 9), np.float32, 61)
    sc = [(3, True, [1, .5]), (1, False, [1, .25, .125]), (0, True, [1, .5]), (2, False, [.5, .5])]
    check_float(oracle, a, sc)


# ---- edge cases -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 8, 9, 16, 29])
def test_orders(oracle, order):
    a = rand_image((50, 700), np.float32, 70 + order)
    coeff = [1.0] + [0.9 / order / (1 + 0.1 * k) for k in range(order)]
    check_float(oracle, a, [(0, True, coeff), (1, False, coeff)], "clamp")


@pytest.mark.parametrize("order", [9, 12, 16, 20, 29])
def test_high_orders_on_long_lines_take_the_warp_chain(oracle, order):
    """Orders above 8 on lines of more than 512 tiles (apps/audio/audio_filter_high_order.cpp sweeps 1..29): the
    generic engine's segmented chain, one warp per (segment, line), with the same-dimension residual between a causal
    and an anticausal scan; float against the fp64 oracle, integers bit exact."""
    coeff = [1.0] + [0.9 / order / (1 + 0.1 * k) for k in range(order)]
    a = (rand_image((1, 70001), np.float32, 300 + order) * 2 - 1).astype(np.float32)
    check_float(oracle, a, [(0, True, coeff)])
    b = (rand_image((3, 40000), np.float32, 320 + order) * 2 - 1).astype(np.float32)
    check_float(oracle, b, [(0, True, coeff), (0, False, coeff)], "clamp")
    ci = [1] + [(-1) ** k * (1 + k % 3) for k in range(order)]                 # integer ring: any coefficients are exact
    c = rand_image((2, 50000), np.uint32, 340 + order)
    check_int(oracle, c, [(0, True, ci), (0, False, ci)])
    # lines along the strided dimension (image mode): 600 tiles of 64 rows
    d = (rand_image((38400, 5), np.float32, 360 + order) * 2 - 1).astype(np.float32)
    check_float(oracle, d, [(1, True, coeff), (1, False, coeff)])


def test_mixed_orders_in_one_dim(oracle):
    a = rand_image((130, 140), np.float32, 80)
    sc = [(0, True, [1, .5]), (0, False, [1, .3, .2, .1]), (1, True, [.5, .2, .2]), (1, True, [1, .4])]
    check_float(oracle, a, sc)
    check_float(oracle, a, sc, "clamp")


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_and_tiles(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    h, w = int(rng.integers(1, 150)), int(rng.integers(1, 150))
    tile = [int(rng.integers(1, 65)), int(rng.integers(1, 65))]
    a = rand_image((h, w), np.float32, seed)
    nsc = int(rng.integers(1, 6))
    sc = []
    for _ in range(nsc):
        r = int(rng.integers(1, 4))
        sc.append((int(rng.integers(0, 2)), bool(rng.integers(0, 2)), [0.4] + [0.5 / r] * r))
    border = "clamp" if seed % 2 else "zero"
    check_float(oracle, a, sc, border, tile=tile, honor_tile=True)
    ai = rand_image((h, w), np.int32, seed)
    sci = [(d, c, [int(rng.integers(-3, 4)) for _ in co]) for d, c, co in sc]
    check_int(oracle, ai, sci, border, tile=tile, honor_tile=True)


def test_int_types(oracle):
    for dt in (np.int32, np.uint16, np.uint8, np.int8):
        a = rand_image((70, 90), dt, 90)
        check_int(oracle, a, [(0, True, [1, 1]), (1, False, [2, -1, 3]), (0, False, [1, 1])])


def test_no_scans_is_copy_and_empty_input():
    a = rand_image((10, 12), np.float32, 91)
    np.testing.assert_array_equal(run_gpu(a, []), a)
    e = np.zeros((0, 7), np.float32)
    assert run_gpu(e, [(0, True, [1, 1])]).shape == (0, 7)


def test_device_resident_execute_in_and_out_of_place(oracle):
    import torch
    a = rand_image((200, 300), np.float32, 92)
    sc = [(0, True, G3), (0, False, G3), (1, True, G3), (1, False, G3)]
    plan = Plan((300, 200), "f32", [Scan(*s) for s in sc], "clamp")
    src = torch.from_numpy(a).cuda()
    dst = plan.execute(src)
    torch.cuda.synchronize()
    host = plan.realize(a)
    np.testing.assert_array_equal(dst.cpu().numpy(), host)
    plan.execute(src, src)                      # in place
    torch.cuda.synchronize()
    np.testing.assert_array_equal(src.cpu().numpy(), host)
    assert plan.num_launches >= 2 and "pass" in plan.describe()
