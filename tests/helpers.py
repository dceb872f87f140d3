"""Shared test helpers: literal pure-Python recurrence (App. A of SURVEY.md) and error metrics."""
import numpy as np


def literal_scan_line(f, coeff, causal, clamp, dtype):
    """The recurrence of RecFilter::add_filter on ONE line, written straight from
    /root/reference/lib/recfilter.cpp:306-343 (independent of oracle.c)."""
    n = len(f)
    T = np.dtype(dtype).type
    is_int = np.issubdtype(np.dtype(dtype), np.integer)
    bits = np.dtype(dtype).itemsize * 8
    # add_filter takes vector<float>; integer types truncate the coefficient and wrap
    c = [np.array(int(x) & ((1 << bits) - 1), dtype=np.uint64).astype(dtype)[()] if is_int else T(np.float32(x))
         for x in coeff]
    f = np.array(f, dtype=dtype)
    with np.errstate(over="ignore"):
        for s in range(n):
            i = s if causal else n - 1 - s
            acc = T(c[0] * f[i])
            for j in range(1, len(c)):
                idx = i - j if causal else i + j
                if 0 <= idx < n:
                    tap = f[idx]
                elif clamp:
                    tap = f[0] if causal else f[n - 1]      # in-place array: f[i] itself when s == 0
                else:
                    tap = T(0)
                acc = T(acc + T(c[j] * tap))
            f[i] = acc
    return f


def literal_filter(a, scans, border="zero"):
    """Apply scans [(dim, causal, coeff)] in order; dim 0 = last numpy axis."""
    a = np.array(a, copy=True)
    nd = a.ndim
    for dim, causal, coeff in scans:
        ax = nd - 1 - dim
        a = np.apply_along_axis(literal_scan_line, ax, a, coeff, causal, border == "clamp", a.dtype)
    return a


def rel_err(out, ref):
    """max |out-ref| / max|ref| (scale-relative) and the reference's own metric
    100*|ref-out|/(ref+1e-9) % (lib/recfilter.h:818-825)."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    return float(np.abs(out - ref).max() / scale)


def rand_image(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    if np.issubdtype(np.dtype(dtype), np.integer):
        info = np.iinfo(dtype)
        return rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    return rng.random(size=shape, dtype=np.float32).astype(dtype)


def ref_metric_percent(out, ref):
    """The reference's own error metric, lib/recfilter.h:818-825: max over samples of 100 * |ref - out| / (ref + 1e-9)
    (no abs on ref: meant for positive images) -- a PER-SAMPLE relative error in percent."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float((100.0 * np.abs(ref - out) / (ref + 1e-9)).max())
