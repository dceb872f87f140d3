"""IIR coefficient design, host side (restates /root/reference/lib/iir_coeff.cpp).

Same free functions, same argument meaning and the same float32 rounding points as
the reference so that the coefficient sets agree to the last ulp or two:

    gaussian_weights(sigma, order)     lib/iir_coeff.cpp:162-177 (van Vliet-Young-Verbeek
                                       poles rescaled to sigma, :38-159)
    integral_image_coeff(n)            lib/iir_coeff.cpp:222-234
    overlap_feedback_coeff(a, b)       lib/iir_coeff.cpp:236-263
    gaussian_box_filter(k, sigma)      lib/iir_coeff.cpp:205-220

All return ``{b0, a1..ar}`` lists as passed to add_filter: feedback terms are ADDED.
"""
from __future__ import annotations

import cmath
import math

import numpy as np

f32 = np.float32


def _qs(s: f32) -> f32:
    # double arithmetic, float result (lib/iir_coeff.cpp:38-40)
    return f32(0.00399341 + 0.4715161 * float(s))


def _ds_real(d: f32, s: f32) -> f32:
    return f32(math.pow(float(d), 1.0 / float(_qs(s))))


def _ds_complex(d: complex, s: f32) -> complex:
    q = float(_qs(s))
    return cmath.rect(math.pow(abs(d), 1.0 / q), cmath.phase(d) / q)


def _weights1(s: f32):
    d3 = f32(1.86543)
    d = _ds_real(d3, s)
    b0 = f32(-(1.0 - float(d)) / float(d))
    a1 = f32(-1.0 / float(d))
    return b0, a1


def _weights2(s: f32):
    d = _ds_complex(complex(1.41650, 1.00829), s)
    n2 = f32(abs(d))
    n2 = f32(n2 * n2)
    re = f32(d.real)
    b0 = f32((1.0 - 2.0 * float(re) + float(n2)) / float(n2))
    a1 = f32(-2.0 * float(re) / float(n2))
    a2 = f32(1.0 / float(n2))
    return b0, a1, a2


def _weights3(s: f32):
    b10, a11 = _weights1(s)
    b20, a21, a22 = _weights2(s)
    a1 = f32(a11 + a21)
    a2 = f32(f32(a11 * a21) + a22)
    a3 = f32(a11 * a22)
    b0 = f32(b10 * b20)
    return b0, a1, a2, a3


def gaussian_weights(sigma: float, order: int) -> list[float]:
    """Feed-forward + feedback coefficients of the order-1/2/3 recursive Gaussian."""
    s = f32(sigma)
    a = [f32(0.0)] * (order + 1)
    if order == 1:
        w = _weights1(s)
    elif order == 2:
        w = _weights2(s)
    else:
        w = _weights3(s)
    for i, v in enumerate(w):
        if i < len(a):
            a[i] = v
    return [float(a[0])] + [float(f32(-v)) for v in a[1:]]


def _binomial_coeff(n: int, i: int, r: float) -> f32:
    n_choose_i = math.factorial(n) // (math.factorial(i) * math.factorial(n - i))
    return f32(math.pow(-r, i) * float(f32(n_choose_i)))


def integral_image_coeff(n: int) -> list[float]:
    """n-fold integral: feedback = -binomial expansion of (1-x)^n, e.g. {1,1}, {1,2,-1}."""
    c = [1.0]
    for i in range(1, n + 1):
        c.append(float(f32(-1.0) * _binomial_coeff(n, i, 1.0)))
    return c


def overlap_feedback_coeff(a: list[float], b: list[float]) -> list[float]:
    """Feedback of the cascade of two all-pole filters (polynomial product)."""
    aa = [f32(1.0)] + [f32(-x) for x in a]
    bb = [f32(1.0)] + [f32(-x) for x in b]
    c = [f32(0.0)] * (len(aa) + len(bb) - 1)
    for i in range(len(c)):
        for j in range(i + 1):
            if j < len(aa) and i - j < len(bb):
                c[i] = f32(c[i] + f32(aa[j] * bb[i - j]))
    return [float(f32(-x)) for x in c[1:]]


def gaussian_box_filter(k: int, sigma: float) -> int:
    """Box width so that k box iterations approximate a Gaussian of std sigma."""
    total = f32(0.0)
    alpha = f32(0.005)
    limit = int(math.floor((float(k) - 1.0) / 2.0))
    for i in range(limit + 1):
        f_k, f_i = math.factorial(k), math.factorial(i)
        f_k_i, f_k_1 = math.factorial(k - i), math.factorial(k - 1)
        f = f32(f_k // (f_i * f_k_i))
        p = f32(math.pow(-1.0, i) / float(f32(f_k_1)))
        total = f32(float(total) + float(p) * float(f) * math.pow(float(f32(k)) / 2.0 - i, k - 1))
    total = f32(math.sqrt(2.0 * math.pi) * float(f32(total + alpha)) * float(f32(sigma)))
    return int(math.ceil(float(total)))
