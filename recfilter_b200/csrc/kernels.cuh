#pragma once
/*
 * kernels.cu -- hand-written sm_100a kernels of the tiled recursive-filter engine.
 *
 * Replaces, as one family of kernels, what the reference obtains by rewriting a
 * Halide Func into the tiled DAG and JIT-compiling it through libnvvm
 * (/root/reference/lib/split.cpp:503-665 intra-tile term, :256-499 tail extraction,
 * :743-867 inter-tile carry completion, :1215-1633 cross-dimension residual,
 * :1008-1130 + :1647-1780 final fix-up; schedules lib/recfilter.cpp:682-870).
 *
 *   tile_kernel<TAILS>   K1: scan every tile with zero history, emit per-scan tails
 *   chain_*_kernel       K2: tail -> carry completion along one dimension
 *   cross_kernel         K3: residual of completed x carries on the d tails
 *   tile_kernel<FINAL>   K4: re-scan every tile from its completed carries, write result
 *
 * A tile is TILE x TILE samples handled by one CTA of TILE threads.  Row scans
 * (contiguous dimension) run thread-per-row after a shared-memory transposition
 * (128-bit loads from a 68-word padded row: conflict free); column scans run
 * thread-per-column straight out of shared memory / global memory (coalesced).
 * Each thread keeps its whole line segment in registers, so all scans of a
 * dimension are applied back to back without touching memory.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "engine.h"

namespace rfb {

constexpr int SROW = TILE + 4;   // padded shared-memory row (words): 16B aligned, conflict free

// ---------------------------------------------------------------------------------------------
// arithmetic in the compute type
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float    madd(float a, float b, float c)          { return fmaf(a, b, c); }
__device__ __forceinline__ uint32_t madd(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
__device__ __forceinline__ double   madd(double a, double b, double c)       { return fma(a, b, c); }


/*
 * One scan of a register-resident line segment.
 *   v      the TILE samples of this thread (only [0,len) are meaningful)
 *   h      history, h[0] = most recent output in scan order; on return the tail
 *   c      {b0, a1..aR} (orders below R are zero padded)
 *   clampb this tile starts at a closed, clamped image border: the first step
 *          reads the not-yet-updated border sample for every tap, later steps the
 *          updated one (/root/reference/lib/recfilter.cpp:330-336)
 * The feed-forward product and the far taps are accumulated first and the
 * nearest tap last, so consecutive samples are one FMA apart on the critical path.
 */
template <typename CT, int R, bool CAUSAL, bool FULL>
__device__ __forceinline__ void scan_regs(CT (&v)[TILE], CT (&h)[R], const CT (&c)[R + 1],
                                          const int len, const bool clampb)
{
#pragma unroll
    for (int p = 0; p < TILE; ++p) {
        const int i = CAUSAL ? p : TILE - 1 - p;
        if (FULL || i < len) {
            const CT x = v[i];
            const bool first = CAUSAL ? (i == 0) : (FULL ? (i == TILE - 1) : (i == len - 1));
            CT acc = c[0] * x;
            if (first && clampb) {
#pragma unroll
                for (int k = R; k >= 1; --k) acc = madd(c[k], x, acc);
#pragma unroll
                for (int k = 0; k < R; ++k) h[k] = acc;
            } else {
#pragma unroll
                for (int k = R; k >= 1; --k) acc = madd(c[k], h[k - 1], acc);
#pragma unroll
                for (int k = R - 1; k >= 1; --k) h[k] = h[k - 1];
                h[0] = acc;
            }
            v[i] = acc;
        }
    }
}

// High orders (R > 8) are rare (the audio order sweep): keep the code small with rolled loops;
// the line segment then lives in local memory.
template <typename CT, int R>
__device__ __noinline__ void scan_rolled(CT (&v)[TILE], CT (&h)[R], const CT (&c)[R + 1],
                                         const int len, const bool clampb, const bool causal)
{
#pragma unroll 1
    for (int p = 0; p < len; ++p) {
        const int i = causal ? p : len - 1 - p;
        const CT x = v[i];
        CT acc = c[0] * x;
        if (p == 0 && clampb) {
#pragma unroll 1
            for (int k = R; k >= 1; --k) acc = madd(c[k], x, acc);
#pragma unroll 1
            for (int k = 0; k < R; ++k) h[k] = acc;
        } else {
#pragma unroll 1
            for (int k = R; k >= 1; --k) acc = madd(c[k], h[k - 1], acc);
#pragma unroll 1
            for (int k = R - 1; k >= 1; --k) h[k] = h[k - 1];
            h[0] = acc;
        }
        v[i] = acc;
    }
}

template <typename CT, int R>
__device__ __forceinline__ void scan_dispatch(CT (&v)[TILE], CT (&h)[R], const CT (&c)[R + 1],
                                              const int len, const bool clampb, const bool causal,
                                              const bool full)
{
    if constexpr (R > 8) {
        // full tiles: the unrolled register scan (history shifts become register renaming); with the rolled loops the
        // history lives in local memory and every sample costs ~3R local accesses (measured: 16.7 M samples of order
        // 9..15 took 0.9 ms in K1 alone, order 17..29 2.0 ms)
        if (full) {
            if (causal) scan_regs<CT, R, true,  true>(v, h, c, len, clampb);
            else        scan_regs<CT, R, false, true>(v, h, c, len, clampb);
        } else {
            // local copies: the arrays handed to the rolled (noinline) scan live in local memory, the caller's stay in registers
            CT w[TILE], hh[R], cc[R + 1];
#pragma unroll
            for (int i = 0; i < TILE; ++i) w[i] = v[i];
#pragma unroll
            for (int k = 0; k < R; ++k) hh[k] = h[k];
#pragma unroll
            for (int k = 0; k <= R; ++k) cc[k] = c[k];
            scan_rolled<CT, R>(w, hh, cc, len, clampb, causal);
#pragma unroll
            for (int i = 0; i < TILE; ++i) v[i] = w[i];
#pragma unroll
            for (int k = 0; k < R; ++k) h[k] = hh[k];
        }
    } else {
        if (full) {
            if (causal) scan_regs<CT, R, true,  true>(v, h, c, len, clampb);
            else        scan_regs<CT, R, false, true>(v, h, c, len, clampb);
        } else {
            // ragged / small tiles (image edges, tiny test shapes): rolled loops on a local copy so
            // that v itself stays in registers on the hot path
            CT w[TILE];
#pragma unroll
            for (int i = 0; i < TILE; ++i) w[i] = v[i];
            scan_rolled<CT, R>(w, h, c, len, clampb, causal);
#pragma unroll
            for (int i = 0; i < TILE; ++i) v[i] = w[i];
        }
    }
}

enum { MODE_TAILS = 0, MODE_FINAL = 1 };

// ---------------------------------------------------------------------------------------------
// K1 / K4: the tile kernel
// ---------------------------------------------------------------------------------------------
template <typename CT, int R, int MODE>
__global__ void __launch_bounds__(TILE)
tile_kernel(const __grid_constant__ PassParams<CT, R> p, const CT* in, CT* out)
{
    __shared__ __align__(16) CT tile[TILE][SROW];

    const int tid = threadIdx.x;

    // ---- which tile ----
    int bx, bd; int64_t o;          // image mode
    int64_t g = 0, line = 0;        // signal mode
    if (p.signal_mode) {
        const int64_t ngroups = (p.nbx + TILE - 1) / TILE;
        g    = blockIdx.x % ngroups;
        line = blockIdx.x / ngroups;
        bx = 0; bd = 0; o = 0;
    } else {
        int64_t b = blockIdx.x;
        bx = (int)(b % p.nbx); b /= p.nbx;
        bd = (int)(b % p.nbd); o = b / p.nbd;
    }

    const int64_t Nx = p.Nx;
    // image-mode tile extents
    const int cols_valid = p.signal_mode ? 0 : (int)min((int64_t)p.tx, Nx - (int64_t)bx * p.tx);
    const int rows_valid = p.signal_mode ? 0 : (int)min((int64_t)p.td, p.Nd - (int64_t)bd * p.td);
    const int64_t tile_base = p.signal_mode ? line * Nx
                                            : (o * p.Nd + (int64_t)bd * p.td) * Nx + (int64_t)bx * p.tx;

    CT v[TILE];

    const bool has_x = p.mx > 0;

    if (has_x) {
        // ---- stage the tile in shared memory (coalesced) ----
        const bool vec_ok = ((Nx & 3) == 0) && ((p.tx & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        if (vec_ok) {
            // 16 threads x 16 B cover one row; 4 rows per step
            const int c4 = tid & 15;
            const int rsub = tid >> 4;
#pragma unroll 4
            for (int r0 = 0; r0 < TILE; r0 += 4) {
                const int r = r0 + rsub;
                int64_t rbase; int rlen;
                if (p.signal_mode) {
                    const int64_t tb = g * TILE + r;
                    rbase = tile_base + tb * p.tx;
                    rlen  = (tb < p.nbx) ? (int)min((int64_t)p.tx, Nx - tb * p.tx) : 0;
                } else {
                    rbase = tile_base + (int64_t)r * Nx;
                    rlen  = (r < rows_valid) ? cols_valid : 0;
                }
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (c4 * 4 < rlen) q = *reinterpret_cast<const uint4*>(in + rbase + c4 * 4);
                *reinterpret_cast<uint4*>(&tile[r][c4 * 4]) = q;
            }
        } else {
            for (int r = 0; r < TILE; ++r) {
                int64_t rbase; int rlen;
                if (p.signal_mode) {
                    const int64_t tb = g * TILE + r;
                    rbase = tile_base + tb * p.tx;
                    rlen  = (tb < p.nbx) ? (int)min((int64_t)p.tx, Nx - tb * p.tx) : 0;
                } else {
                    rbase = tile_base + (int64_t)r * Nx;
                    rlen  = (r < rows_valid) ? cols_valid : 0;
                }
                tile[r][tid] = (tid < rlen) ? in[rbase + tid] : (CT)0;
            }
        }
        __syncthreads();

        // ---- row scans: thread tid owns row tid ----
        int my_len; int64_t my_bx; int64_t lx; bool row_ok;
        if (p.signal_mode) {
            my_bx  = g * TILE + tid;
            row_ok = my_bx < p.nbx;
            my_len = row_ok ? (int)min((int64_t)p.tx, Nx - my_bx * p.tx) : 0;
            lx     = line;
        } else {
            my_bx  = bx;
            row_ok = tid < rows_valid;
            my_len = cols_valid;
            lx     = o * p.Nd + (int64_t)bd * p.td + tid;
        }
        if (row_ok) {
#pragma unroll
            for (int c4 = 0; c4 < TILE / 4; ++c4) {
                const uint4 q = *reinterpret_cast<const uint4*>(&tile[tid][c4 * 4]);
                v[c4 * 4 + 0] = *reinterpret_cast<const CT*>(&q.x);
                v[c4 * 4 + 1] = *reinterpret_cast<const CT*>(&q.y);
                v[c4 * 4 + 2] = *reinterpret_cast<const CT*>(&q.z);
                v[c4 * 4 + 3] = *reinterpret_cast<const CT*>(&q.w);
            }
            const bool full = (my_len == TILE);
            for (int s = 0; s < p.mx; ++s) {
                const bool causal = p.sx.causal[s] != 0;
                const bool closed = causal ? (my_bx == 0 && p.x_lo_closed)
                                           : (my_bx == p.nbx - 1 && p.x_hi_closed);
                const int64_t idx0 = p.signal_mode
                    ? ((int64_t)s * R * p.nlx + lx) * p.nbx + my_bx
                    : ((int64_t)s * R * p.nbx + my_bx) * p.nlx + lx;
                const int64_t kstride = (int64_t)p.nbx * p.nlx;
                CT h[R];
#pragma unroll
                for (int k = 0; k < R; ++k)
                    h[k] = (MODE == MODE_FINAL && !closed) ? p.CX[idx0 + k * kstride] : (CT)0;
                scan_dispatch<CT, R>(v, h, p.sx.coef[s], my_len, closed && p.clamp, causal, full);
                if (MODE == MODE_TAILS) {
#pragma unroll
                    for (int k = 0; k < R; ++k) p.TX[idx0 + k * kstride] = h[k];
                }
            }
        }
        if (MODE == MODE_TAILS && p.md == 0) return;

        // ---- write the row back for the column phase / the coalesced store ----
        // (each thread only ever touched its own row since the staging barrier)
        if (row_ok) {
#pragma unroll
            for (int c4 = 0; c4 < TILE / 4; ++c4) {
                uint4 q;
                *reinterpret_cast<CT*>(&q.x) = v[c4 * 4 + 0];
                *reinterpret_cast<CT*>(&q.y) = v[c4 * 4 + 1];
                *reinterpret_cast<CT*>(&q.z) = v[c4 * 4 + 2];
                *reinterpret_cast<CT*>(&q.w) = v[c4 * 4 + 3];
                *reinterpret_cast<uint4*>(&tile[tid][c4 * 4]) = q;
            }
        }
        __syncthreads();

        if (p.md == 0) {
            // x-only pass, FINAL: cooperative coalesced store
            if (vec_ok) {
                const int c4 = tid & 15;
                const int rsub = tid >> 4;
#pragma unroll 4
                for (int r0 = 0; r0 < TILE; r0 += 4) {
                    const int r = r0 + rsub;
                    int64_t rbase; int rlen;
                    if (p.signal_mode) {
                        const int64_t tb = g * TILE + r;
                        rbase = tile_base + tb * p.tx;
                        rlen  = (tb < p.nbx) ? (int)min((int64_t)p.tx, Nx - tb * p.tx) : 0;
                    } else {
                        rbase = tile_base + (int64_t)r * Nx;
                        rlen  = (r < rows_valid) ? cols_valid : 0;
                    }
                    if (c4 * 4 < rlen)
                        *reinterpret_cast<uint4*>(out + rbase + c4 * 4) =
                            *reinterpret_cast<const uint4*>(&tile[r][c4 * 4]);
                }
            } else {
                for (int r = 0; r < TILE; ++r) {
                    int64_t rbase; int rlen;
                    if (p.signal_mode) {
                        const int64_t tb = g * TILE + r;
                        rbase = tile_base + tb * p.tx;
                        rlen  = (tb < p.nbx) ? (int)min((int64_t)p.tx, Nx - tb * p.tx) : 0;
                    } else {
                        rbase = tile_base + (int64_t)r * Nx;
                        rlen  = (r < rows_valid) ? cols_valid : 0;
                    }
                    if (tid < rlen) out[rbase + tid] = tile[r][tid];
                }
            }
            return;
        }
    }

    // ---- column scans: thread tid owns column tid (image mode only) ----
    const bool col_ok = tid < cols_valid;
    if (!col_ok) return;
    if (has_x) {
#pragma unroll
        for (int i = 0; i < TILE; ++i) v[i] = tile[i][tid];
    } else {
#pragma unroll
        for (int i = 0; i < TILE; ++i)
            v[i] = (i < rows_valid) ? in[tile_base + (int64_t)i * Nx + tid] : (CT)0;
    }
    {
        const bool full = (rows_valid == TILE);
        const int64_t ly = o * Nx + (int64_t)bx * p.tx + tid;
        const int64_t kstride = (int64_t)p.nbd * p.nly;
        for (int s = 0; s < p.md; ++s) {
            const bool causal = p.sd.causal[s] != 0;
            const bool closed = causal ? (bd == 0 && p.d_lo_closed)
                                       : (bd == p.nbd - 1 && p.d_hi_closed);
            const int64_t idx0 = ((int64_t)s * R * p.nbd + bd) * p.nly + ly;
            CT h[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
                h[k] = (MODE == MODE_FINAL && !closed) ? p.CY[idx0 + k * kstride] : (CT)0;
            scan_dispatch<CT, R>(v, h, p.sd.coef[s], rows_valid, closed && p.clamp, causal, full);
            if (MODE == MODE_TAILS) {
#pragma unroll
                for (int k = 0; k < R; ++k) p.TY[idx0 + k * kstride] = h[k];
            }
        }
    }
    if (MODE == MODE_FINAL) {
#pragma unroll
        for (int i = 0; i < TILE; ++i)
            if (i < rows_valid) out[tile_base + (int64_t)i * Nx + tid] = v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// K2: carry completion along one dimension for one scan.
//   c[j]   = history entering tile j = completed tail of the previous tile in scan order
//   tail'  = T[s][j] + sum_{q<s} M[var(j)][q][s] * C[q][j]      (same-dimension residual,
//            /root/reference/lib/split.cpp:912-1004)
//   tau[j] = tail' + P[var(j)][s] * c[j]                        (lib/split.cpp:832-846)
// Tiles are grouped in segments so that long signals are chained in parallel:
// local (zero carry per segment) -> top (over segments) -> fix (add P^k * segment carry).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_variant(int j, int nb)
{
    if (nb == 1) return V_SINGLE;
    if (j == 0) return V_FIRST;
    if (j == nb - 1) return V_LAST;
    return V_INTERIOR;
}

template <typename TT, int R>
__device__ __forceinline__ void matvec_acc(TT (&y)[R], const TT* __restrict__ m, const TT (&x)[R])
{
#pragma unroll (R <= 8 ? R : 1)
    for (int k = 0; k < R; ++k) {
        TT a = y[k];
#pragma unroll (R <= 8 ? R : 1)
        for (int kk = 0; kk < R; ++kk) a = madd(m[k * R + kk], x[kk], a);
        y[k] = a;
    }
}

template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_local_kernel(const __grid_constant__ ChainParams<CT, R> p)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.nl * p.nseg) return;
    const int64_t l = gid % p.nl;
    const int g = (int)(gid / p.nl);          // segment index in SCAN order
    const int s = p.s;
    typedef typename TabType<CT>::type TT;

    TT tau[R];
#pragma unroll (R <= 8 ? R : 1)
    for (int k = 0; k < R; ++k) tau[k] = (g == 0 && p.ext) ? (TT)p.ext[(int64_t)k * p.nl + l] : (TT)0;

    const int j_begin = g * p.seg;
    const int j_end = min(p.nb, j_begin + p.seg);
    for (int jj = j_begin; jj < j_end; ++jj) {
        const int j = p.causal ? jj : p.nb - 1 - jj;      // tile index in memory
        const int var = tile_variant(j, p.nb);
        const int64_t base = (int64_t)j * p.tile_stride + l * p.line_stride;
        TT t[R];
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) {
            const int64_t idx = ((int64_t)s * R + k) * p.plane + base;
            p.C[idx] = (CT)tau[k];
            t[k] = (TT)p.T[idx];
        }
        for (int q = 0; q < s; ++q) {
            TT cq[R];
#pragma unroll (R <= 8 ? R : 1)
            for (int k = 0; k < R; ++k) cq[k] = (TT)p.C[((int64_t)q * R + k) * p.plane + base];
            matvec_acc<TT, R>(t, p.M + (((int64_t)var * p.S + q) * p.S + s) * R * R, cq);
        }
        matvec_acc<TT, R>(t, p.P + ((int64_t)var * p.S + s) * R * R, tau);
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) tau[k] = t[k];
    }
    if (p.nseg > 1) {
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) p.SEGT[((int64_t)k * p.nseg + g) * p.nl + l] = tau[k];
    } else if (p.tail_out) {
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) p.tail_out[(int64_t)k * p.nl + l] = (CT)tau[k];
    }
}

// serial over segments, one thread per line
template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_top_kernel(const __grid_constant__ ChainParams<CT, R> p)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= p.nl) return;
    typedef typename TabType<CT>::type TT;
    TT sigma[R];
#pragma unroll (R <= 8 ? R : 1)
    for (int k = 0; k < R; ++k) sigma[k] = (TT)0;   // segment 0 already started from ext
    for (int g = 0; g < p.nseg; ++g) {
        TT t[R];
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) {
            const int64_t idx = ((int64_t)k * p.nseg + g) * p.nl + l;
            p.SEGC[idx] = sigma[k];
            t[k] = p.SEGT[idx];
        }
        const TT* Pm = p.Pseg + (g == p.nseg - 1 ? R * R : 0);
        matvec_acc<TT, R>(t, Pm, sigma);
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) sigma[k] = t[k];
    }
    if (p.tail_out) {
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) p.tail_out[(int64_t)k * p.nl + l] = (CT)sigma[k];
    }
}

// add the propagated segment carry to the provisional carries of a segment
template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_fix_kernel(const __grid_constant__ ChainParams<CT, R> p)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.nl * (p.nseg - 1)) return;
    const int64_t l = gid % p.nl;
    const int g = (int)(gid / p.nl) + 1;
    const int s = p.s;
    typedef typename TabType<CT>::type TT;
    TT u[R];
#pragma unroll (R <= 8 ? R : 1)
    for (int k = 0; k < R; ++k) u[k] = p.SEGC[((int64_t)k * p.nseg + g) * p.nl + l];
    const int j_begin = g * p.seg;
    const int j_end = min(p.nb, j_begin + p.seg);
    for (int jj = j_begin; jj < j_end; ++jj) {
        const int j = p.causal ? jj : p.nb - 1 - jj;
        const int var = tile_variant(j, p.nb);
        const int64_t base = (int64_t)j * p.tile_stride + l * p.line_stride;
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) {
            const int64_t idx = ((int64_t)s * R + k) * p.plane + base;
            p.C[idx] = (CT)((TT)p.C[idx] + u[k]);
        }
        TT t[R];
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) t[k] = (TT)0;
        matvec_acc<TT, R>(t, p.P + ((int64_t)var * p.S + s) * R * R, u);
#pragma unroll (R <= 8 ? R : 1)
        for (int k = 0; k < R; ++k) u[k] = t[k];
    }
}

// ---------------------------------------------------------------------------------------------
// K2 for high orders (R >= 16): the same three steps with one WARP per (segment, line) -- lane k owns component k of
// the history, a matrix-vector product is R shuffles + R FMAs per lane with the matrix row read conflict-free from a
// transposed copy in shared memory (interior tiles; the border variants, a handful of tiles, read global memory).
// One thread per line, as above, keeps R-element arrays that it indexes in rolled loops, i.e. in local memory: a
// 16.7 M-sample order-9..15 signal spent 25 ms of its 26.7 ms in the chain, order 17..29 90 of 93 ms
// (scripts/high_order_time.py).
// ---------------------------------------------------------------------------------------------
template <typename TT, int R>
__device__ __forceinline__ TT wmatvec(const TT* __restrict__ m, const bool transposed, const TT x, const int kl)
{
    TT a[4] = { (TT)0, (TT)0, (TT)0, (TT)0 };       // four partial sums: the FMA chain is R/4 deep
#pragma unroll
    for (int kk = 0; kk < R; kk += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const TT xe = __shfl_sync(0xffffffffu, x, kk + e);
            a[e] = madd(transposed ? m[(kk + e) * R + kl] : m[kl * R + kk + e], xe, a[e]);
        }
    }
    return (a[0] + a[1]) + (a[2] + a[3]);
}
// stage `count` R x R matrices, transposed, into shared memory (the whole block takes part)
template <typename TT, int R>
__device__ __forceinline__ void wstage(TT* dst, const TT* __restrict__ src, int count)
{
    for (int i = threadIdx.x; i < count * R * R; i += blockDim.x) {
        const int mtx = i / (R * R), e = i - mtx * R * R, k = e / R, kk = e - k * R;
        dst[mtx * R * R + kk * R + k] = src[(int64_t)mtx * R * R + e];
    }
    __syncthreads();
}
constexpr int WCHAIN_MAXQ = 2;      // same-dimension residual matrices kept in shared memory (earlier scans q < 2)
constexpr int WCHAIN_BATCH = 8;     // chain steps whose (recurrence-independent) loads are issued together

template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_local_wkernel(const __grid_constant__ ChainParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    __shared__ TT sm[(1 + WCHAIN_MAXQ) * R * R];
    const int s = p.s;
    // interior variant: P[s], then M[q -> s] for q < min(s, WCHAIN_MAXQ)
    wstage<TT, R>(sm, p.P + ((int64_t)V_INTERIOR * p.S + s) * R * R, 1);
    for (int q = 0; q < s && q < WCHAIN_MAXQ; ++q)
        wstage<TT, R>(sm + (1 + q) * R * R, p.M + (((int64_t)V_INTERIOR * p.S + q) * p.S + s) * R * R, 1);
    const int lane = threadIdx.x & 31, kl = lane < R ? lane : 0;
    const bool act = lane < R;
    const int64_t gid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= p.nl * p.nseg) return;
    const int64_t l = gid % p.nl;
    const int g = (int)(gid / p.nl);          // segment index in SCAN order
    TT tau = (g == 0 && p.ext && act) ? (TT)p.ext[(int64_t)kl * p.nl + l] : (TT)0;
    const int j_begin = g * p.seg;
    const int j_end = min(p.nb, j_begin + p.seg);
    auto tile_of = [&](int jj) { return p.causal ? jj : p.nb - 1 - jj; };
    auto index = [&](int sc, int j) { return ((int64_t)sc * R + kl) * p.plane + (int64_t)j * p.tile_stride + l * p.line_stride; };
    // the tails do not depend on the recurrence: WB steps' worth are asked for at once, ahead of the serial part
    // (one step ahead left ~500 cycles of L2 latency exposed in every step)
    for (int jb = j_begin; jb < j_end; jb += WCHAIN_BATCH) {
        TT tb[WCHAIN_BATCH];
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) tb[e] = (jb + e < j_end) ? (TT)p.T[index(s, tile_of(jb + e))] : (TT)0;
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) {
            const int jj = jb + e;
            if (jj < j_end) {
                const int j = tile_of(jj);
                const int var = tile_variant(j, p.nb);
                TT t = tb[e];
                if (act) p.C[index(s, j)] = (CT)tau;
                for (int q = 0; q < s; ++q) {
                    const TT cq = (TT)p.C[index(q, j)];
                    const bool fast = var == V_INTERIOR && q < WCHAIN_MAXQ;
                    t = t + wmatvec<TT, R>(fast ? sm + (1 + q) * R * R : p.M + (((int64_t)var * p.S + q) * p.S + s) * R * R, fast, cq, kl);
                }
                const bool fast = var == V_INTERIOR;
                t = t + wmatvec<TT, R>(fast ? sm : p.P + ((int64_t)var * p.S + s) * R * R, fast, tau, kl);
                tau = t;
            }
        }
    }
    if (!act) return;
    if (p.nseg > 1) p.SEGT[((int64_t)kl * p.nseg + g) * p.nl + l] = tau;
    else if (p.tail_out) p.tail_out[(int64_t)kl * p.nl + l] = (CT)tau;
}

template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_top_wkernel(const __grid_constant__ ChainParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    __shared__ TT sm[2 * R * R];
    wstage<TT, R>(sm, p.Pseg, 2);
    const int lane = threadIdx.x & 31, kl = lane < R ? lane : 0;
    const bool act = lane < R;
    const int64_t l = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= p.nl) return;
    TT sigma = (TT)0;                             // segment 0 already started from ext
    for (int gb = 0; gb < p.nseg; gb += WCHAIN_BATCH) {
        TT tb[WCHAIN_BATCH];
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) tb[e] = (gb + e < p.nseg) ? p.SEGT[((int64_t)kl * p.nseg + gb + e) * p.nl + l] : (TT)0;
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) {
            const int g = gb + e;
            if (g < p.nseg) {
                if (act) p.SEGC[((int64_t)kl * p.nseg + g) * p.nl + l] = sigma;
                sigma = tb[e] + wmatvec<TT, R>(sm + (g == p.nseg - 1 ? R * R : 0), true, sigma, kl);
            }
        }
    }
    if (p.tail_out && act) p.tail_out[(int64_t)kl * p.nl + l] = (CT)sigma;
}

template <typename CT, int R>
__global__ void __launch_bounds__(128)
chain_fix_wkernel(const __grid_constant__ ChainParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    __shared__ TT sm[R * R];
    const int s = p.s;
    wstage<TT, R>(sm, p.P + ((int64_t)V_INTERIOR * p.S + s) * R * R, 1);
    const int lane = threadIdx.x & 31, kl = lane < R ? lane : 0;
    const bool act = lane < R;
    const int64_t gid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= p.nl * (p.nseg - 1)) return;
    const int64_t l = gid % p.nl;
    const int g = (int)(gid / p.nl) + 1;
    TT u = p.SEGC[((int64_t)kl * p.nseg + g) * p.nl + l];
    const int j_begin = g * p.seg;
    const int j_end = min(p.nb, j_begin + p.seg);
    for (int jb = j_begin; jb < j_end; jb += WCHAIN_BATCH) {
        TT cb[WCHAIN_BATCH];
        int64_t ib[WCHAIN_BATCH];
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) {
            const int jj = min(jb + e, j_end - 1);
            const int j = p.causal ? jj : p.nb - 1 - jj;
            ib[e] = ((int64_t)s * R + kl) * p.plane + (int64_t)j * p.tile_stride + l * p.line_stride;
            cb[e] = (TT)p.C[ib[e]];
        }
#pragma unroll
        for (int e = 0; e < WCHAIN_BATCH; ++e) {
            const int jj = jb + e;
            if (jj < j_end) {
                const int j = p.causal ? jj : p.nb - 1 - jj;
                const int var = tile_variant(j, p.nb);
                if (act) p.C[ib[e]] = (CT)(cb[e] + u);
                const bool fast = var == V_INTERIOR;
                u = wmatvec<TT, R>(fast ? sm : p.P + ((int64_t)var * p.S + s) * R * R, fast, u, kl);
            }
        }
    }
}

template <typename CT, int R>
__global__ void __launch_bounds__(TILE)
cross_kernel(const __grid_constant__ CrossParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TT* A = reinterpret_cast<TT*>(smem_raw);            // [md][mx][R][R]
    TT* red = A + p.md * p.mx * R * R;                  // [TILE/32] scratch

    const int tid = threadIdx.x;
    int64_t b = blockIdx.x;
    const int bx = (int)(b % p.nbx); b /= p.nbx;
    const int bd = (int)(b % p.nbd);
    const int64_t o = b / p.nbd;
    const int cols_valid = (int)min((int64_t)p.tx, p.Nx - (int64_t)bx * p.tx);
    const int rows_valid = (int)min((int64_t)p.td, p.Nd - (int64_t)bd * p.td);
    const int vx = tile_variant(bx, p.nbx);
    const int vd = tile_variant(bd, p.nbd);
    const int64_t lx = o * p.Nd + (int64_t)bd * p.td + tid;
    const int64_t ly = o * p.Nx + (int64_t)bx * p.tx + tid;
    const int64_t kstride_x = (int64_t)p.nbx * p.nlx;
    const int64_t kstride_d = (int64_t)p.nbd * p.nly;

    // phase 1: A[s][q][k][kk] = sum_rows L[vd][s][k][row] * CX[q][kk][bx][row]
    for (int q = 0; q < p.mx; ++q) {
        TT cx[R];
#pragma unroll (R <= 8 ? R : 1)
        for (int kk = 0; kk < R; ++kk)
            cx[kk] = (tid < rows_valid) ? (TT)p.CX[((int64_t)q * R * p.nbx + bx) * p.nlx + lx + kk * kstride_x] : (TT)0;
        for (int s = 0; s < p.md; ++s) {
#pragma unroll (R <= 8 ? R : 1)
            for (int k = 0; k < R; ++k) {
                const TT lv = (tid < rows_valid) ? p.L[(((int64_t)vd * p.md + s) * R + k) * TILE + tid] : (TT)0;
#pragma unroll (R <= 8 ? R : 1)
                for (int kk = 0; kk < R; ++kk) {
                    TT part = lv * cx[kk];
                    // block reduction (2 warps)
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1)
                        part = part + __shfl_xor_sync(0xffffffffu, part, off);
                    if ((tid & 31) == 0) red[tid >> 5] = part;
                    __syncthreads();
                    if (tid == 0) {
                        TT tot = (TT)0;
                        for (int w = 0; w < TILE / 32; ++w) tot = tot + red[w];
                        A[((s * p.mx + q) * R + k) * R + kk] = tot;
                    }
                    __syncthreads();
                }
            }
        }
    }
    __syncthreads();

    // phase 2: column tid
    if (tid < cols_valid) {
        for (int s = 0; s < p.md; ++s) {
            TT d[R];
#pragma unroll (R <= 8 ? R : 1)
            for (int k = 0; k < R; ++k) d[k] = (TT)0;
            for (int q = 0; q < p.mx; ++q) {
                TT gq[R];
#pragma unroll (R <= 8 ? R : 1)
                for (int kk = 0; kk < R; ++kk)
                    gq[kk] = p.G[(((int64_t)vx * p.mx + q) * TILE + tid) * R + kk];
                matvec_acc<TT, R>(d, A + (s * p.mx + q) * R * R, gq);
            }
            const int64_t idx0 = ((int64_t)s * R * p.nbd + bd) * p.nly + ly;
#pragma unroll (R <= 8 ? R : 1)
            for (int k = 0; k < R; ++k) p.TY[idx0 + k * kstride_d] = (CT)((TT)p.TY[idx0 + k * kstride_d] + d[k]);
        }
    }
}


} // namespace rfb
