#pragma once
/*
 * fused.cuh -- the fast path: fused two-dimension tile kernels with register-resident
 * lines and a parallel fp64 carry algebra.
 *
 * What it replaces in the reference (/root/reference): the Halide-generated stages of
 * lib/split.cpp -- intra-tile term :503-665, tail extraction :256-499, inter-tile carry
 * completion :743-867 (+ same-dimension residual :912-1004), cross-dimension residual
 * :1215-1633, final term :1008-1130 / :1647-1780 -- and their GPU schedules
 * (lib/recfilter.cpp:682-870).
 *
 * A pass sees the array as [No][Nd][Nx] (Nx contiguous).  A tile is TS x TS samples and
 * is handled by one CTA of TS threads:
 *
 *   load          one thread issues TS/32 TMA box loads (cp.async.bulk.tensor, 128-byte
 *                 swizzle) onto an mbarrier: no per-thread address arithmetic, no LDG;
 *   column phase  thread t owns column t: TS conflict-free shared loads into registers,
 *                 every scan along d is run on the register line, line written back;
 *   row phase     thread t owns row t (128-bit shared loads through the swizzle), every
 *                 scan along x on the register line;
 *   store         (pass 2 only) rows -> shared, fence.proxy.async, TMA box stores.
 *
 * Scans along different dimensions commute exactly (lib/split.cpp:207-213), so running
 * the d scans before the x scans is legal; it makes the d carries "pure" (no
 * cross-dimension term).
 *
 * Every scan is run with unit feed-forward (y' = x + sum a_k y'[i-k]); the product of the
 * feed-forward coefficients is applied once per sample when pass 2 writes the result.
 *
 *   fused_tile_kernel<P1>   zero-history scans, emits per-scan tails (TY, TX)
 *   fchain_kernel           tails -> carries along one dimension, all scans, segmented
 *   fcross_kernel           A = L_x * CY  per tile (cross-dimension residual, part 1)
 *   fchain_kernel (x)       adds G_y * A to the x tails on the fly (part 2), then chains
 *   fused_tile_kernel<P2>   re-scan from the completed carries, scale, store
 */
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include "engine.h"
#include "fused_params.h"

namespace rfb {

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float    fmadd(float a, float b, float c)          { return fmaf(a, b, c); }
__device__ __forceinline__ uint32_t fmadd(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
__device__ __forceinline__ double   fmadd(double a, double b, double c)       { return fma(a, b, c); }

// streaming 4-byte load: read-only path, do not allocate in L1 (the tile is touched once)
template <typename CT>
__device__ __forceinline__ CT ld_stream(const CT* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return *reinterpret_cast<CT*>(&r);
}
__device__ __forceinline__ uint4 ld_stream4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

/*
 * One unit-feed-forward scan of a register line of N samples.
 *   h       history on entry (h[0] most recent), tail on return
 *   a       a[0] clamp factor, a[1..R] feedback
 *   clampb  the line starts at a closed, clamped image border: the first sample sees
 *           itself (divided by b0 in the scaled domain) on every tap, later samples the
 *           updated border sample (/root/reference/lib/recfilter.cpp:330-336)
 * The far taps are accumulated first so that consecutive samples are one FMA apart.
 */
template <typename CT, int R, int N, bool CAUSAL>
__device__ __forceinline__ void scan_line(CT (&v)[N], CT (&h)[R], const CT (&a)[R + 1], const bool clampb)
{
    if (clampb) {
        const CT x0 = v[CAUSAL ? 0 : N - 1] * a[0];
#pragma unroll
        for (int k = 0; k < R; ++k) h[k] = x0;
    }
#pragma unroll
    for (int p = 0; p < N; ++p) {
        const int i = CAUSAL ? p : N - 1 - p;
        CT acc = v[i];
#pragma unroll
        for (int k = R; k >= 1; --k) acc = fmadd(a[k], h[k - 1], acc);
#pragma unroll
        for (int k = R - 1; k >= 1; --k) h[k] = h[k - 1];
        h[0] = acc;
        if (p == 0 && clampb) {
#pragma unroll
            for (int k = 1; k < R; ++k) h[k] = acc;
        }
        v[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG / UTMASTG / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tile load: box (x fastest) -> shared memory, completion counted on the mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int x, int y, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, int x, int y, const void* smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// P1 / P2: the tile kernel.
// Shared memory holds the tile as TS/32 TMA boxes of [TS rows][32 columns] (128 B rows) in the
// 128-byte swizzle: the 16-byte chunk c of row r lives at chunk c ^ (r & 7).  With that layout the
// thread-per-column accesses (32 lanes = the 32 words of one row) and the thread-per-row 128-bit
// accesses (8 lanes = 8 different chunks) are both bank-conflict free, without padding.
// ---------------------------------------------------------------------------------------------
template <typename CT, int R, int TS, int MODE>
__global__ void __launch_bounds__(TS, (TS == 128 ? 3 : 6))
fused_tile_kernel(const __grid_constant__ FusedParams<CT, R> p, const __grid_constant__ CUtensorMap tm_in,
                  const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int NBOX = TS / 32;
    constexpr int BOX_BYTES = TS * 128;
    extern __shared__ __align__(16) unsigned char fsmem_raw[];
    // the swizzle is a function of the shared-memory address: boxes must start on a 1024 B boundary
    unsigned char* tile = fsmem_raw + ((1024u - (smem_u32(fsmem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + NBOX * BOX_BYTES);

    const int tid = threadIdx.x;
    int64_t b = p.reverse ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x;
    const int bx = (int)(b % p.nbx); b /= p.nbx;
    const int bd = (int)(b % p.nbd);
    const int64_t o = b / p.nbd;
    const int x0 = bx * TS;
    const int y0 = (int)(o * p.Nd + (int64_t)bd * TS);                  // row of the [No*Nd][Nx] matrix

    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, x0 + bb * 32, y0, bar);
    }

    CT v[TS];
    CT hn[R];                                                          // history of the next scan (prefetched)

    if (p.md > 0) {
        // ---- column phase: thread tid owns column tid ----
        const int64_t ly = o * p.Nx + (int64_t)bx * TS + tid;
        const int64_t kstride = (int64_t)p.nbd * p.nly;
        auto load_cy = [&](int s) {
            const bool causal = p.sd.causal[s] != 0;
            const bool closed = causal ? (bd == 0 && p.d_lo_closed) : (bd == p.nbd - 1 && p.d_hi_closed);
            const int64_t idx0 = ((int64_t)s * R * p.nbd + bd) * p.nly + ly;
#pragma unroll
            for (int k = 0; k < R; ++k) hn[k] = (MODE == FMODE_P2 && !closed) ? p.CY[idx0 + k * kstride] : (CT)0;
        };
        load_cy(0);                                                    // in flight while the tile lands
        const uint32_t cbase = smem_u32(tile) + (tid >> 5) * BOX_BYTES + (((tid & 31) >> 2) << 4) + ((tid & 3) << 2);
        mbar_wait(bar, 0);
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"((cbase ^ ((i & 7) << 4)) + i * 128));
            v[i] = *reinterpret_cast<CT*>(&w);
        }
        for (int s = 0; s < p.md; ++s) {
            const bool causal = p.sd.causal[s] != 0;
            const bool closed = causal ? (bd == 0 && p.d_lo_closed) : (bd == p.nbd - 1 && p.d_hi_closed);
            CT a[R + 1];
#pragma unroll
            for (int k = 0; k <= R; ++k) a[k] = p.sd.a[s][k];
            CT h[R];
#pragma unroll
            for (int k = 0; k < R; ++k) h[k] = hn[k];
            if (MODE == FMODE_P2 && s + 1 < p.md) load_cy(s + 1);
            if (causal) scan_line<CT, R, TS, true >(v, h, a, closed && p.clamp);
            else        scan_line<CT, R, TS, false>(v, h, a, closed && p.clamp);
            if (MODE == FMODE_P1) {
                const int64_t idx0 = ((int64_t)s * R * p.nbd + bd) * p.nly + ly;
#pragma unroll
                for (int k = 0; k < R; ++k) p.TY[idx0 + k * kstride] = h[k];
            }
        }
        if (MODE == FMODE_P1 && (p.mx == 0 || p.nbx == 1)) return;      // no x tails wanted
        // ---- back to shared memory (same words this thread read) ----
        const CT g = (MODE == FMODE_P2 && p.mx == 0) ? p.gain : (CT)1;
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            const CT w = (MODE == FMODE_P2 && p.mx == 0) ? v[i] * g : v[i];
            asm volatile("st.shared.b32 [%0], %1;" :: "r"((cbase ^ ((i & 7) << 4)) + i * 128),
                         "r"(*reinterpret_cast<const uint32_t*>(&w)) : "memory");
        }
    }

    if (p.mx > 0) {
        // ---- row phase: thread tid owns row tid ----
        const int64_t lx = o * p.Nd + (int64_t)bd * TS + tid;
        const int64_t kstride = (int64_t)p.nbx * p.nlx;
        auto load_cx = [&](int s) {
            const bool causal = p.sx.causal[s] != 0;
            const bool closed = causal ? (bx == 0 && p.x_lo_closed) : (bx == p.nbx - 1 && p.x_hi_closed);
            const int64_t idx0 = ((int64_t)s * R * p.nbx + bx) * p.nlx + lx;
#pragma unroll
            for (int k = 0; k < R; ++k) hn[k] = (MODE == FMODE_P2 && !closed) ? p.CX[idx0 + k * kstride] : (CT)0;
        };
        load_cx(0);
        if (p.md > 0) __syncthreads(); else mbar_wait(bar, 0);
        const uint32_t rbase = smem_u32(tile) + tid * 128;
        const uint32_t rx = (tid & 7) << 4;
#pragma unroll
        for (int c4 = 0; c4 < TS / 4; ++c4) {
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(rbase + (c4 >> 3) * BOX_BYTES + ((((c4 & 7) << 4)) ^ rx)));
            v[c4 * 4 + 0] = *reinterpret_cast<const CT*>(&q.x);
            v[c4 * 4 + 1] = *reinterpret_cast<const CT*>(&q.y);
            v[c4 * 4 + 2] = *reinterpret_cast<const CT*>(&q.z);
            v[c4 * 4 + 3] = *reinterpret_cast<const CT*>(&q.w);
        }
        for (int s = 0; s < p.mx; ++s) {
            const bool causal = p.sx.causal[s] != 0;
            const bool closed = causal ? (bx == 0 && p.x_lo_closed) : (bx == p.nbx - 1 && p.x_hi_closed);
            CT a[R + 1];
#pragma unroll
            for (int k = 0; k <= R; ++k) a[k] = p.sx.a[s][k];
            CT h[R];
#pragma unroll
            for (int k = 0; k < R; ++k) h[k] = hn[k];
            if (MODE == FMODE_P2 && s + 1 < p.mx) load_cx(s + 1);
            if (causal) scan_line<CT, R, TS, true >(v, h, a, closed && p.clamp);
            else        scan_line<CT, R, TS, false>(v, h, a, closed && p.clamp);
            if (MODE == FMODE_P1) {
                const int64_t idx0 = ((int64_t)s * R * p.nbx + bx) * p.nlx + lx;
#pragma unroll
                for (int k = 0; k < R; ++k) p.TX[idx0 + k * kstride] = h[k];
            }
        }
        if constexpr (MODE == FMODE_P2) {
            // scale, rows back to shared memory (each thread only touches its own row)
#pragma unroll
            for (int c4 = 0; c4 < TS / 4; ++c4) {
                uint4 q;
                *reinterpret_cast<CT*>(&q.x) = v[c4 * 4 + 0] * p.gain;
                *reinterpret_cast<CT*>(&q.y) = v[c4 * 4 + 1] * p.gain;
                *reinterpret_cast<CT*>(&q.z) = v[c4 * 4 + 2] * p.gain;
                *reinterpret_cast<CT*>(&q.w) = v[c4 * 4 + 3] * p.gain;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};"
                             :: "r"(rbase + (c4 >> 3) * BOX_BYTES + ((((c4 & 7) << 4)) ^ rx)),
                                "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
            }
        }
    }
    if constexpr (MODE == FMODE_P2) {
        // ---- store: generic-proxy writes -> async proxy, then one thread issues the TMA stores ----
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int bb = 0; bb < NBOX; ++bb) tma_store_2d(&tm_out, x0 + bb * 32, y0, tile + bb * BOX_BYTES);
            tma_store_commit_and_wait_read();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// carry chain along one dimension, all scans of that dimension in one launch
//   tail'  = T[s][j] (+ G_y * A, the cross-dimension residual) + sum_{q<s} M[q->s] * c_q[j]
//   tau[j] = tail' + P[s] * c_s[j];   c_s[next tile in scan order] = tau[j]
// blockDim = (32 lines, nseg segments of FCHAIN_L tiles).  A thread owns FCHAIN_L memory-adjacent
// tiles of one line for every scan (so it can re-read the carries of earlier scans it wrote
// itself); segment tails are exchanged through shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ftile_variant(int j, int nb)
{
    if (nb == 1) return V_SINGLE;
    if (j == 0) return V_FIRST;
    if (j == nb - 1) return V_LAST;
    return V_INTERIOR;
}

template <typename TT, int R>
__device__ __forceinline__ void fmatvec_acc(TT (&y)[R], const TT* __restrict__ m, const TT (&x)[R])
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        TT acc = y[k];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) acc = fmadd(__ldg(m + k * R + kk), x[kk], acc);
        y[k] = acc;
    }
}

template <typename CT, int R>
__global__ void __launch_bounds__(32 * 16)
fchain_kernel(const __grid_constant__ FChainParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    extern __shared__ __align__(16) unsigned char fchain_smem[];
    TT* segtail = reinterpret_cast<TT*>(fchain_smem);       // [nseg][R][32]

    const int lane = threadIdx.x, g = threadIdx.y;
    const int64_t l = (int64_t)blockIdx.x * 32 + lane;
    const bool valid = l < p.nl;
    const int64_t lc = valid ? l : p.nl - 1;                 // clamp: keep the barriers uniform
    const int j0 = g * FCHAIN_L;
    const int j1 = min(p.nb, j0 + FCHAIN_L);
    const int cnt = j1 - j0;
    const int64_t plane = (int64_t)p.nb * p.nl;

    // cross-dimension residual: row of G for this line (x chain of a fused pass)
    TT gy[FMAX_SCANS][R];
    int64_t a_base = 0;
    if (p.A) {
        const int64_t o = lc / p.Nd;
        const int64_t rem = lc - o * p.Nd;
        const int bd = (int)(rem / p.TS), i = (int)(rem - (int64_t)bd * p.TS);
        const int vd = ftile_variant(bd, p.nbd);
#pragma unroll
        for (int sd = 0; sd < FMAX_SCANS; ++sd)
#pragma unroll
            for (int k = 0; k < R; ++k)
                gy[sd][k] = sd < p.Sd ? __ldg(p.G + (((int64_t)vd * p.Sd + sd) * p.TS + i) * R + k) : (TT)0;
        a_base = (o * p.nbd + bd) * (int64_t)p.nb;           // tile (o, bd, bx=0)
    }

    for (int s = 0; s < p.S; ++s) {
        const bool causal = p.causal[s] != 0;
        const int gs = causal ? g : p.nseg - 1 - g;          // segment index in scan order
        TT cl[FCHAIN_L][R];
        TT tau[R];
#pragma unroll
        for (int k = 0; k < R; ++k)
            tau[k] = (gs == 0 && p.ext) ? (TT)p.ext[((int64_t)s * R + k) * p.nl + lc] : (TT)0;

#pragma unroll
        for (int t = 0; t < FCHAIN_L; ++t) {
            if (t < cnt) {
                const int j = causal ? j0 + t : j1 - 1 - t;
                const int var = ftile_variant(j, p.nb);
                const int64_t base = (int64_t)j * p.nl + lc;
                TT tt[R];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    cl[t][k] = tau[k];
                    tt[k] = (TT)p.T[((int64_t)s * R + k) * plane + base];
                }
                if (p.A) {
                    const TT* At = p.A + (((a_base + j) * p.Sd) * p.S + s) * (R * R);
                    for (int sd = 0; sd < p.Sd; ++sd) {
                        const TT* Am = At + (int64_t)sd * p.S * (R * R);
#pragma unroll
                        for (int kx = 0; kx < R; ++kx) {
                            TT acc = tt[kx];
#pragma unroll
                            for (int k = 0; k < R; ++k) acc = fmadd(gy[sd][k], __ldg(Am + k * R + kx), acc);
                            tt[kx] = acc;
                        }
                    }
                }
                for (int q = 0; q < s; ++q) {
                    TT cq[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) cq[k] = (TT)p.C[((int64_t)q * R + k) * plane + base];
                    fmatvec_acc<TT, R>(tt, p.M + (((int64_t)var * p.S + q) * p.S + s) * R * R, cq);
                }
                fmatvec_acc<TT, R>(tt, p.P + ((int64_t)var * p.S + s) * R * R, tau);
#pragma unroll
                for (int k = 0; k < R; ++k) tau[k] = tt[k];
            }
        }
        // ---- exchange segment tails ----
#pragma unroll
        for (int k = 0; k < R; ++k) segtail[(gs * R + k) * 32 + lane] = tau[k];
        __syncthreads();
        TT u[R];
#pragma unroll
        for (int k = 0; k < R; ++k) u[k] = (TT)0;
        for (int g2 = 0; g2 < gs; ++g2) {
            TT nu[R];
#pragma unroll
            for (int k = 0; k < R; ++k) nu[k] = segtail[(g2 * R + k) * 32 + lane];
            fmatvec_acc<TT, R>(nu, p.Pseg + ((int64_t)s * p.nseg + g2) * R * R, u);
#pragma unroll
            for (int k = 0; k < R; ++k) u[k] = nu[k];
        }
        if (p.tail_out && gs == p.nseg - 1 && valid) {
            TT fin[R];
#pragma unroll
            for (int k = 0; k < R; ++k) fin[k] = tau[k];
            fmatvec_acc<TT, R>(fin, p.Pseg + ((int64_t)s * p.nseg + gs) * R * R, u);
#pragma unroll
            for (int k = 0; k < R; ++k) p.tail_out[((int64_t)s * R + k) * p.nl + l] = (CT)fin[k];
        }
        // ---- add the propagated segment carry, store the carries ----
#pragma unroll
        for (int t = 0; t < FCHAIN_L; ++t) {
            if (t < cnt) {
                const int j = causal ? j0 + t : j1 - 1 - t;
                const int var = ftile_variant(j, p.nb);
                const int64_t base = (int64_t)j * p.nl + lc;
                if (valid) {
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        p.C[((int64_t)s * R + k) * plane + base] = (CT)(cl[t][k] + u[k]);
                }
                TT nu[R];
#pragma unroll
                for (int k = 0; k < R; ++k) nu[k] = (TT)0;
                fmatvec_acc<TT, R>(nu, p.P + ((int64_t)var * p.S + s) * R * R, u);
#pragma unroll
                for (int k = 0; k < R; ++k) u[k] = nu[k];
            }
        }
        __syncthreads();                                      // segtail is reused by the next scan
    }
}

// ---------------------------------------------------------------------------------------------
// cross-dimension residual, part 1:  A[tile][sd][sx] = sum over the tile's columns of
//   CY_sd[k][col] * L_sx[kx][col]      (R x R per scan pair; one warp per tile)
// The completed d carries change the d-filtered tile by G_y * CY (rows x cols); the x tails P1
// computed on the incomplete tile therefore miss  (G_y * CY) * L_x^T = G_y * A
// (/root/reference/lib/split.cpp:1215-1633).
// ---------------------------------------------------------------------------------------------
template <typename CT, int R, int TS>
__global__ void __launch_bounds__(128)
fcross_kernel(const __grid_constant__ FCrossParams<CT, R> p)
{
    typedef typename TabType<CT>::type TT;
    constexpr int CPL = TS / 32;                 // columns per lane
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (int64_t)p.nbx * p.nbd * p.No;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= ntiles) return;
    int64_t b = w;
    const int bx = (int)(b % p.nbx); b /= p.nbx;
    const int bd = (int)(b % p.nbd);
    const int64_t o = b / p.nbd;
    const int vx = ftile_variant(bx, p.nbx);
    const int64_t ly0 = o * p.Nx + (int64_t)bx * TS + lane;
    const int64_t kstride = (int64_t)p.nbd * p.nly;

    for (int sd = 0; sd < p.Sd; ++sd) {
        TT cy[R][CPL];
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                cy[k][c] = (TT)p.CY[((int64_t)sd * R * p.nbd + bd) * p.nly + k * kstride + ly0 + c * 32];
        for (int q = 0; q < p.Sx; ++q) {
            TT acc[R][R];
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int kx = 0; kx < R; ++kx) acc[k][kx] = (TT)0;
#pragma unroll
            for (int kx = 0; kx < R; ++kx)
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const TT lv = __ldg(p.L + (((int64_t)vx * p.Sx + q) * R + kx) * TS + c * 32 + lane);
#pragma unroll
                    for (int k = 0; k < R; ++k) acc[k][kx] = fmadd(cy[k][c], lv, acc[k][kx]);
                }
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int kx = 0; kx < R; ++kx) {
                    TT x = acc[k][kx];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, off);
                    acc[k][kx] = x;
                }
            if (lane == 0) {
                TT* dst = p.A + ((w * p.Sd + sd) * p.Sx + q) * (R * R);
#pragma unroll
                for (int k = 0; k < R; ++k)
#pragma unroll
                    for (int kx = 0; kx < R; ++kx) dst[k * R + kx] = acc[k][kx];
            }
        }
    }
}

} // namespace rfb
