/*
 * oracle.c -- CPU restatement of RecFilter::add_filter semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under recfilter_b200/ (the product) may
 * include, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg
 * as the checker and the timed CPU baseline.
 *
 * What it restates (all citations into /root/reference):
 *   lib/recfilter.cpp:260-392  RecFilter::add_filter -- one scan is, along one
 *       dimension, in place, over the whole array:
 *         f[i] <- T(b0)*f[i] + T(a1)*tap(1) + ... + T(ar)*tap(r)
 *       summed left to right (:324-341); causal scans walk i = 0..N-1, anticausal
 *       scans walk i = N-1..0 (:308-319);
 *       zero border  : tap(j) = (j <= steps_done) ? f[i -/+ j] : 0   (:338-339)
 *       clamp border : tap(j) = f[clamp(i -/+ j, 0, N-1)]            (:330-336)
 *       on the in-place array (so at the first step every tap reads the not yet
 *       updated border sample, afterwards the updated one).
 *   lib/recfilter.cpp:324,335,338  coefficients are floats cast to the element
 *       type T; integer T truncates them and integer arithmetic wraps.
 *   tests/test_generic_xyz.cpp:56-116, apps/summed_table/summed_table.cpp:67-82,
 *   apps/bspline/bicubic_filter.cpp:124-156 are the reference's own inline loops
 *   of exactly this recurrence; oracle/Makefile compiles those programs unchanged
 *   (_ref/pin/<prog>) and pin_reference.py runs them against this file.
 *
 * Parity status: pinned.  oracle/pin_reference.py (results in tests/golden/PIN_REPORT.json):
 *   zero border, ff = 1 : nine programs of tests/*.cpp, bit exact in the test loops' summation order;
 *   clamped border      : apps/bspline/bicubic_filter (unit gain {1+a,-a}: error 0) and
 *                         apps/bspline/biquintic_cascaded_filter ({1+a,-a,0.1}: gain 1.1, order 2; its check
 *                         applies the scans in another, commuting, order: 4.3e-5 % = fp32 rounding), at widths
 *                         64 and 256; biquintic_overlapped_filter dies in gpu_auto_schedule() with the
 *                         reference's own assertion (lib/recfilter.cpp:699-704), as it does on the reference.
 *   Every pin uses the reference's all-ones input (lib/recfilter.h:695-696); the Gaussian / audio / box apps
 *   carry no check in the reference.
 *
 * Layout: dense, dimension 0 contiguous (Halide::Image layout,
 * lib/recfilter.cpp:970-981).  Arithmetic is done in T with separate multiply
 * and add (compile with -ffp-contract=off) exactly like the serial loops.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_ORDER 64

/*
 * Summation order inside one sample:
 *   0  library order: ((b0*x + a1*t1) + a2*t2) + ...          lib/recfilter.cpp:324-341 (the default)
 *   1  test-loop order: b0*x + ((a1*t1 + a2*t2) + ...)        how the reference's test programs write
 *      their inline checks (tests/test_causal_xy.cpp:57-64: ref += t1 + t2 + t3).  Used only by the
 *      pin (oracle/pin_reference.py) to show that, rounding order aside, the oracle and those loops
 *      agree bit for bit.
 */
static int g_sum_order = 0;
void oracle_set_sum_order(int order) { g_sum_order = order ? 1 : 0; }

enum { ORACLE_F32 = 0, ORACLE_F64 = 1, ORACLE_I32 = 2, ORACLE_U32 = 3,
       ORACLE_I16 = 4, ORACLE_U16 = 5, ORACLE_I8 = 6, ORACLE_U8 = 7 };

/* float -> T conversion of a coefficient (Cast::make(type, coeff)) */
static inline float    cvt_f32(float c) { return c; }
static inline double   cvt_f64(float c) { return (double)c; }
static inline uint32_t cvt_u32(float c) { return (uint32_t)(int64_t)c; }
static inline uint16_t cvt_u16(float c) { return (uint16_t)(int64_t)c; }
static inline uint8_t  cvt_u8 (float c) { return (uint8_t)(int64_t)c; }


/* products in T (integers wrap; done in unsigned to avoid signed-overflow UB) */
static inline float    mul_f32(float a, float b)       { return a * b; }
static inline double   mul_f64(double a, double b)     { return a * b; }
static inline uint32_t mul_u32(uint32_t a, uint32_t b) { return a * b; }
static inline uint16_t mul_u16(uint16_t a, uint16_t b) { return (uint16_t)((uint32_t)a * (uint32_t)b); }
static inline uint8_t  mul_u8 (uint8_t a, uint8_t b)   { return (uint8_t)((uint32_t)a * (uint32_t)b); }

/*
 * One scan over an array viewed as [outer][n][inner] (inner contiguous).
 * Rows along the scanned axis are visited in scan order; for every row the
 * `inner` contiguous samples are independent lines, which is both the literal
 * per-line recurrence and a vectorisable loop.
 */
#define DEFINE_SCAN(NAME, T, CVT, MUL)                                                 \
static void NAME(T* data, int64_t outer, int64_t n, int64_t inner, int causal,      \
                 int clamp, const float* coeff, int ncoeff, int nthreads)           \
{                                                                                   \
    const int r = ncoeff - 1;                                                       \
    T c[ORACLE_MAX_ORDER + 1];                                                      \
    for (int j = 0; j <= r; j++) c[j] = CVT(coeff[j]);                              \
    /* split the independent lines into blocks for OpenMP */                        \
    const int64_t blk = inner >= 64 ? 64 : inner;                                   \
    const int64_t nblk = (inner + blk - 1) / blk;                                   \
    (void)nthreads;                                                                 \
    _Pragma("omp parallel for collapse(2) schedule(static) num_threads(nthreads)")  \
    for (int64_t o = 0; o < outer; o++) {                                           \
        for (int64_t b = 0; b < nblk; b++) {                                        \
            const int64_t x0 = b * blk;                                             \
            const int64_t x1 = x0 + blk < inner ? x0 + blk : inner;                 \
            T* base = data + o * n * inner;                                         \
            for (int64_t s = 0; s < n; s++) {                                       \
                const int64_t i = causal ? s : n - 1 - s;                           \
                T* row = base + i * inner;                                          \
                if (clamp && s == 0) {                                              \
                    /* every tap is the not-yet-updated border sample */            \
                    for (int64_t x = x0; x < x1; x++) {                             \
                        const T old = row[x];                                       \
                        T acc = MUL(c[0], old);                                    \
                        if (g_sum_order) {                                          \
                            T ts = MUL(c[1], old);                                 \
                            for (int j = 2; j <= r; j++) ts = (T)(ts + MUL(c[j], old)); \
                            acc = (T)(acc + ts);                                    \
                        } else                                                      \
                        for (int j = 1; j <= r; j++) acc = (T)(acc + MUL(c[j], old)); \
                        row[x] = acc;                                               \
                    }                                                               \
                    continue;                                                       \
                }                                                                   \
                const T* tap[ORACLE_MAX_ORDER + 1];                                 \
                for (int j = 1; j <= r; j++) {                                      \
                    int64_t sj = s - j;            /* steps back in scan order */   \
                    if (sj < 0) { tap[j] = clamp ? (causal ? base : base + (n - 1) * inner) : 0; } \
                    else        { tap[j] = base + (causal ? sj : n - 1 - sj) * inner; } \
                }                                                                   \
                if (g_sum_order) {                                                  \
                    for (int64_t x = x0; x < x1; x++) {                             \
                        T ts = MUL(c[1], tap[1] ? tap[1][x] : (T)0);               \
                        for (int j = 2; j <= r; j++)                                \
                            ts = (T)(ts + MUL(c[j], tap[j] ? tap[j][x] : (T)0));   \
                        row[x] = (T)(MUL(c[0], row[x]) + ts);                      \
                    }                                                               \
                } else                                                              \
                for (int64_t x = x0; x < x1; x++) {                                 \
                    T acc = MUL(c[0], row[x]);                                     \
                    for (int j = 1; j <= r; j++) {                                  \
                        const T t = tap[j] ? tap[j][x] : (T)0;                      \
                        acc = (T)(acc + MUL(c[j], t));                             \
                    }                                                               \
                    row[x] = acc;                                                   \
                }                                                                   \
            }                                                                       \
        }                                                                           \
    }                                                                               \
}

DEFINE_SCAN(scan_f32, float,    cvt_f32, mul_f32)
DEFINE_SCAN(scan_f64, double,   cvt_f64, mul_f64)
DEFINE_SCAN(scan_u32, uint32_t, cvt_u32, mul_u32)
DEFINE_SCAN(scan_u16, uint16_t, cvt_u16, mul_u16)
DEFINE_SCAN(scan_u8,  uint8_t,  cvt_u8, mul_u8)

/*
 * Contiguous-axis scan (dimension 0): one line per row, serial recurrence with
 * a rolling history -- the literal loop of the reference tests.
 */
#define DEFINE_SCAN_X(NAME, T, CVT, MUL)                                               \
static void NAME(T* data, int64_t lines, int64_t n, int causal, int clamp,          \
                 const float* coeff, int ncoeff, int nthreads)                      \
{                                                                                   \
    const int r = ncoeff - 1;                                                       \
    T c[ORACLE_MAX_ORDER + 1];                                                      \
    for (int j = 0; j <= r; j++) c[j] = CVT(coeff[j]);                              \
    (void)nthreads;                                                                 \
    _Pragma("omp parallel for schedule(static) num_threads(nthreads)")              \
    for (int64_t l = 0; l < lines; l++) {                                           \
        T* f = data + l * n;                                                        \
        for (int64_t s = 0; s < n; s++) {                                           \
            const int64_t i = causal ? s : n - 1 - s;                               \
            const T cur = f[i];                                                     \
            T acc = MUL(c[0], cur);                                                \
            T ts = (T)0;                                                            \
            for (int j = 1; j <= r; j++) {                                          \
                T t;                                                                \
                if (s - j >= 0)      t = f[causal ? i - j : i + j];                 \
                else if (clamp)      t = (s == 0) ? cur : f[causal ? 0 : n - 1];    \
                else                 t = (T)0;                                      \
                if (g_sum_order) ts = (j == 1) ? MUL(c[j], t) : (T)(ts + MUL(c[j], t)); \
                else             acc = (T)(acc + MUL(c[j], t));                    \
            }                                                                       \
            f[i] = g_sum_order ? (T)(acc + ts) : acc;                               \
        }                                                                           \
    }                                                                               \
}

DEFINE_SCAN_X(scanx_f32, float,    cvt_f32, mul_f32)
DEFINE_SCAN_X(scanx_f64, double,   cvt_f64, mul_f64)
DEFINE_SCAN_X(scanx_u32, uint32_t, cvt_u32, mul_u32)
DEFINE_SCAN_X(scanx_u16, uint16_t, cvt_u16, mul_u16)
DEFINE_SCAN_X(scanx_u8,  uint8_t,  cvt_u8, mul_u8)

/*
 * Apply one scan in place.
 *   data     dense array, extents[0] contiguous
 *   dtype    ORACLE_*
 *   dim      scanned dimension (0 = contiguous)
 *   causal   1: +x, 0: -x
 *   clamp    0: zero border, 1: clamped border (set_clamped_image_border)
 *   coeff    {b0, a1..ar}, ncoeff = r+1 >= 2   (lib/recfilter.cpp:274-287)
 *   nthreads OpenMP threads over independent lines (1 = the serial loop)
 * returns 0, or -1 on bad arguments.
 */
int oracle_scan(void* data, int dtype, int ndim, const int64_t* extents, int dim,
                int causal, int clamp, const float* coeff, int ncoeff, int nthreads)
{
    if (!data || ndim < 1 || ndim > 8 || dim < 0 || dim >= ndim) return -1;
    if (ncoeff < 2 || ncoeff > ORACLE_MAX_ORDER + 1) return -1;
    if (nthreads < 1) nthreads = 1;
    int64_t inner = 1, outer = 1;
    for (int d = 0; d < dim; d++) inner *= extents[d];
    for (int d = dim + 1; d < ndim; d++) outer *= extents[d];
    const int64_t n = extents[dim];
    if (n <= 0 || inner <= 0 || outer <= 0) return 0;   /* empty: nothing to do */

    if (dim == 0) {
        switch (dtype) {
        case ORACLE_F32: scanx_f32((float*)data, outer, n, causal, clamp, coeff, ncoeff, nthreads); break;
        case ORACLE_F64: scanx_f64((double*)data, outer, n, causal, clamp, coeff, ncoeff, nthreads); break;
        case ORACLE_I32: case ORACLE_U32:
            scanx_u32((uint32_t*)data, outer, n, causal, clamp, coeff, ncoeff, nthreads); break;
        case ORACLE_I16: case ORACLE_U16:
            scanx_u16((uint16_t*)data, outer, n, causal, clamp, coeff, ncoeff, nthreads); break;
        case ORACLE_I8: case ORACLE_U8:
            scanx_u8((uint8_t*)data, outer, n, causal, clamp, coeff, ncoeff, nthreads); break;
        default: return -1;
        }
        return 0;
    }
    switch (dtype) {
    case ORACLE_F32: scan_f32((float*)data, outer, n, inner, causal, clamp, coeff, ncoeff, nthreads); break;
    case ORACLE_F64: scan_f64((double*)data, outer, n, inner, causal, clamp, coeff, ncoeff, nthreads); break;
    case ORACLE_I32: case ORACLE_U32:
        scan_u32((uint32_t*)data, outer, n, inner, causal, clamp, coeff, ncoeff, nthreads); break;
    case ORACLE_I16: case ORACLE_U16:
        scan_u16((uint16_t*)data, outer, n, inner, causal, clamp, coeff, ncoeff, nthreads); break;
    case ORACLE_I8: case ORACLE_U8:
        scan_u8((uint8_t*)data, outer, n, inner, causal, clamp, coeff, ncoeff, nthreads); break;
    default: return -1;
    }
    return 0;
}

/*
 * Apply a whole filter: scans in add_filter order (lib/recfilter.cpp:343,
 * "ONE update definition per scan, applied in add order").
 *   scan_dim[s], scan_causal[s], scan_ncoeff[s]; coefficients are concatenated
 *   in `coeffs`.
 */
int oracle_filter(void* data, int dtype, int ndim, const int64_t* extents, int clamp,
                  int nscans, const int* scan_dim, const int* scan_causal,
                  const int* scan_ncoeff, const float* coeffs, int nthreads)
{
    const float* c = coeffs;
    for (int s = 0; s < nscans; s++) {
        int rc = oracle_scan(data, dtype, ndim, extents, scan_dim[s], scan_causal[s], clamp,
                             c, scan_ncoeff[s], nthreads);
        if (rc) return rc;
        c += scan_ncoeff[s];
    }
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
