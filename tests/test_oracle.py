"""CPU tests of the oracle (oracle/oracle.c): it must equal the literal per-line recurrence of
/root/reference/lib/recfilter.cpp:306-343, scipy's lfilter, and the analytic known answers
SURVEY.md 8(c) lists.  The pin against the reference's own test programs lives in
tests/test_golden.py."""
import numpy as np
import pytest
from scipy.signal import lfilter

from helpers import literal_filter, rand_image
from recfilter_b200.filters import gaussian_weights, integral_image_coeff


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.uint32, np.int16, np.uint8])
@pytest.mark.parametrize("border", ["zero", "clamp"])
def test_oracle_equals_literal_2d(oracle, dtype, border):
    a = rand_image((9, 13), dtype, 1)
    if np.issubdtype(np.dtype(dtype), np.integer):
        scans = [(0, True, [1, 1]), (0, False, [1, 2, -1]), (1, True, [2, 1, -1, 3]), (1, False, [1, 1])]
    else:
        scans = [(0, True, [0.5, 0.25, 0.125]), (0, False, [1.0, 0.5, -0.25, 0.1]),
                 (1, True, [0.7, 0.3]), (1, False, [1.0, 0.4, 0.1])]
    ref = literal_filter(a, scans, border)
    out = oracle.apply_filter(a, scans, border)
    assert out.dtype == a.dtype
    np.testing.assert_array_equal(out, ref)       # bit exact, floats included (same op order)


def test_oracle_equals_literal_3d(oracle):
    a = rand_image((5, 6, 7), np.float32, 2)
    scans = [(0, True, [1, .5, .25]), (0, False, [1, .5, .125]), (1, True, [1, .5, .0625]),
             (1, False, [1, .5, .125]), (2, True, [1, .5, .25]), (2, False, [1, .5, .0625])]
    np.testing.assert_array_equal(oracle.apply_filter(a, scans), literal_filter(a, scans))


def test_oracle_threads_agree(oracle):
    a = rand_image((40, 70), np.float32, 3)
    scans = [(0, True, [1, .5, .25]), (1, False, [1, .5, .125])]
    np.testing.assert_array_equal(oracle.apply_filter(a, scans, threads=1), oracle.apply_filter(a, scans, threads=4))


def test_oracle_vs_scipy_lfilter(oracle):
    # zero border, causal: y = b0 x + sum a_k y[n-k]  <=>  lfilter([b0], [1, -a1, .., -ar])
    x = rand_image((3, 500), np.float64, 4)
    coeff = [float(np.float32(c)) for c in (0.3, 0.9, -0.4, 0.1)]     # add_filter takes floats
    ref = lfilter([coeff[0]], [1.0] + [-c for c in coeff[1:]], x, axis=-1)
    out = oracle.apply_filter(x, [(0, True, coeff)])
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    refa = lfilter([coeff[0]], [1.0] + [-c for c in coeff[1:]], x[:, ::-1], axis=-1)[:, ::-1]
    np.testing.assert_allclose(oracle.apply_filter(x, [(0, False, coeff)]), refa, rtol=1e-12, atol=1e-12)


def test_sat_of_ones_is_analytic(oracle):
    # SAT of ones = (x+1)(y+1)   (apps/summed_table/summed_table.cpp:43-46 with all-ones input)
    h, w = 37, 53
    a = np.ones((h, w), np.uint32)
    out = oracle.apply_filter(a, [(0, True, [1, 1]), (1, True, [1, 1])])
    yy, xx = np.mgrid[0:h, 0:w]
    np.testing.assert_array_equal(out, ((xx + 1) * (yy + 1)).astype(np.uint32))


def test_u32_wraparound(oracle):
    a = np.full((4, 300), 0xF0000000, np.uint32)
    out = oracle.apply_filter(a, [(0, True, [1, 1])])
    ref = (np.cumsum(a.astype(np.uint64), axis=1) & 0xFFFFFFFF).astype(np.uint32)
    np.testing.assert_array_equal(out, ref)


def test_unit_gain_clamped_constant_is_fixed_point(oracle):
    # clamped border + unit DC gain: a constant image stays (numerically) constant
    W3 = gaussian_weights(5.0, 3)
    a = np.full((64, 64), 0.75, np.float32)
    scans = [(0, True, W3), (0, False, W3), (1, True, W3), (1, False, W3)]
    out = oracle.apply_filter(a, scans, "clamp")
    np.testing.assert_allclose(out, 0.75, rtol=2e-4)


def test_known_coefficients():
    # SURVEY.md App. A known answers (tolerance 1e-6: port check, not a bit pin)
    np.testing.assert_allclose(gaussian_weights(5.0, 1), [0.23203820, 0.76796180], atol=1e-6)
    np.testing.assert_allclose(gaussian_weights(5.0, 2), [0.09758424, 1.52838480, -0.62596899], atol=1e-6)
    np.testing.assert_allclose(gaussian_weights(5.0, 3), [0.02264327, 2.29634666, -1.79971004, 0.48072028], atol=1e-6)
    assert integral_image_coeff(1) == [1.0, 1.0]
    assert integral_image_coeff(2) == [1.0, 2.0, -1.0]


def test_empty_input(oracle):
    a = np.zeros((0, 5), np.float32)
    assert oracle.apply_filter(a, [(0, True, [1, 1])]).shape == (0, 5)
