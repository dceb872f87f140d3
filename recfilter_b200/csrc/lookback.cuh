#pragma once
/*
 * lookback.cuh -- single-pass kernels with decoupled look-back for filters whose scans all run one
 * way per dimension (summed-area tables, the box filters' integral images, causal audio filters).
 *
 * What it replaces in the reference (/root/reference): for such filters the reference still builds the
 * full multi-stage pipeline of lib/split.cpp -- intra-tile term :503-665, tails :256-499, the SERIAL
 * inter-tile loop :832-846 (`ctail[t] = tail[t] + R * ctail[t-1]`, scheduled as its own kernels by
 * lib/recfilter.cpp:671-677), final term :1008-1130 -- i.e. the image is read twice and written once
 * (12 B/sample for 4-byte types).  Here a tile is read once and written once (8 B/sample): every CTA
 *
 *   1. takes a ticket (tiles are handed out in scan order, so every tile a CTA will ever wait for is
 *      already running or finished: no deadlock whatever the block scheduler does),
 *   2. loads its tile with TMA and scans it with zero history -> the tile's own tail ("aggregate"),
 *      published with a status word,
 *   3. looks back over its predecessors: a predecessor that already knows its completed tail
 *      ("inclusive") ends the walk, one that only has an aggregate contributes P^j * aggregate and the walk
 *      goes on -- the serial chain of the reference is replaced by a race that ends after a few tiles,
 *   4. publishes its own inclusive tail, re-scans the tile from the carry (the rounding of the serial
 *      recurrence, not of a fix-up sum) and stores it with TMA.
 *
 * The carry algebra runs in the difference basis of fused.cuh (exact for integer rings, so summed-area
 * tables stay bit exact).  For float filters the walk length depends on timing, so the last bits of a
 * result may differ from run to run (as with any decoupled look-back); integer results never do.
 *
 *   lb_tile_kernel     2-D (and stacked 2-D) arrays, at most one scan along x and one along d, orders <= 4:
 *                      the d scan is completed first (look-back up the tile column), the x scan runs on the
 *                      d-complete tile (look-back along the tile row) -> no cross-dimension residual.
 *   lb_signal_kernel   long 1-D signals (rows of 128 samples, 128 rows per CTA), one scan, orders <= 8:
 *                      rows are chained inside the CTA with a Kogge-Stone scan of R-vectors (warp shuffles,
 *                      matrices P^(2^i) as kernel constants), CTAs by a warp-wide look-back window.
 */
#include "fused.cuh"
#include "lookback_params.h"

namespace rfb {

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// payload loads bypass L1 (the line may have been cached before the producer wrote it)
template <typename CT>
__device__ __forceinline__ CT ld_cg(const CT* p)
{
    uint32_t r;
    asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return *reinterpret_cast<CT*>(&r);
}

// wait until a tile of this launch has published something; returns its state (LB_AGGREGATE / LB_INCLUSIVE)
__device__ __forceinline__ uint32_t lb_wait_status(const uint32_t* st, uint32_t epoch, uint32_t* err)
{
    uint32_t spins = 0;
    while (true) {
        const uint32_t s = ld_acquire_u32(st);
        if ((s >> 2) == epoch && (s & 3u) != LB_NONE) return s & 3u;
        // never hang the device: give up after LB_SPIN_LIMIT polls, and at once when another CTA already has
        if (++spins > LB_SPIN_LIMIT || ((spins & 1023u) == 0u && *reinterpret_cast<volatile uint32_t*>(err) != 0u)) {
            *err = 1u;
            return LB_INCLUSIVE;
        }
        if (spins > 8) __nanosleep(40);
    }
}

template <typename CT, int R, int N>
__device__ __forceinline__ void scan_dir(CT (&v)[N], CT (&h)[R], const CT (&a)[R + 1], const bool causal, const bool clampb)
{
    if (causal) scan_line<CT, R, N, true >(v, h, a, clampb);
    else        scan_line<CT, R, N, false>(v, h, a, clampb);
}

// y += M x   (M in kernel-constant or global memory, uniform address)
template <typename CT, int R>
__device__ __forceinline__ void lb_matvec_acc(CT (&y)[R], const CT* m, const CT (&x)[R])
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        CT acc = y[k];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) acc = fmadd(m[k * R + kk], x[kk], acc);
        y[k] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// 2-D single pass
// ---------------------------------------------------------------------------------------------
/*
 * One dimension of a tile: on entry every thread holds its line of the tile in v (the column of thread tid
 * in the d phase, the row in the x phase); on return v is the line filtered with the carry that enters the
 * tile.  `reload` re-reads the line from shared memory (the unfiltered values: needed when the line was first
 * scanned with zero history for the aggregate).
 *   sidx    scan-order index of this tile, pstride the index distance to the previous tile of the line
 *   j       position of the tile along the dimension in scan order (0 = starts at the closed border)
 */
template <typename CT, int R, int TS, typename Reload>
__device__ __forceinline__ void lb_phase(CT (&v)[TS], const LBDim<CT, R>& dm, const int j, const int nb, const uint32_t sidx,
                                         const uint32_t pstride, const uint32_t epoch, const bool clamp, uint32_t* err,
                                         volatile uint32_t* sflag, const int tid, Reload reload)
{
    const bool causal = dm.causal != 0;
    const bool first = j == 0, last = j == nb - 1;
    const uint32_t st_agg = (epoch << 2) | LB_AGGREGATE, st_inc = (epoch << 2) | LB_INCLUSIVE;
    CT a[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) a[k] = dm.a[k];
    CT X[R];                                             // carry entering the tile, difference basis
#pragma unroll
    for (int k = 0; k < R; ++k) X[k] = (CT)0;

    bool have = first;
    if (!first) {
        // the tile before this one may already be complete (usual along d: it started a tile row earlier)
        if (tid == 0) sflag[1] = ld_acquire_u32(dm.status + (sidx - pstride)) == st_inc ? 1u : 0u;
        __syncthreads();
        have = sflag[1] != 0u;
        if (have) {
            const CT* src = dm.inc + (size_t)(sidx - pstride) * R * TS + tid;
#pragma unroll
            for (int k = 0; k < R; ++k) X[k] = ld_cg(src + k * TS);
        }
    }
    if (!have) {
        // aggregate: the tile's own tail (zero history), published before anything is waited for
        CT h[R];
#pragma unroll
        for (int k = 0; k < R; ++k) h[k] = (CT)0;
        scan_dir<CT, R, TS>(v, h, a, causal, false);
        fdiff_fwd<CT, R>(h);
        if (!last) {
            CT* dst = dm.agg + (size_t)sidx * R * TS + tid;
#pragma unroll
            for (int k = 0; k < R; ++k) dst[k * TS] = h[k];
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_u32(dm.status + sidx, st_agg);
        }
        // look back (every warp walks on its own; the status of a tile is one word, so a warp is uniform)
        for (int q = 1; q <= j; ++q) {
            const uint32_t pidx = sidx - (uint32_t)q * pstride;
            const uint32_t state = lb_wait_status(dm.status + pidx, epoch, err);
            const CT* src = (state == LB_INCLUSIVE ? dm.inc : dm.agg) + (size_t)pidx * R * TS + tid;
            CT y[R];
#pragma unroll
            for (int k = 0; k < R; ++k) y[k] = ld_cg(src + k * TS);
            lb_matvec_acc<CT, R>(X, dm.Ppow + (size_t)(q - 1) * R * R, y);
            if (state == LB_INCLUSIVE) break;
        }
        if (!last) {
            // completed tail = aggregate + P * carry: successors need not wait for the re-scan
            lb_matvec_acc<CT, R>(h, dm.P, X);
            CT* dst = dm.inc + (size_t)sidx * R * TS + tid;
#pragma unroll
            for (int k = 0; k < R; ++k) dst[k * TS] = h[k];
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_u32(dm.status + sidx, st_inc);
        }
        reload();
    }
    CT hist[R];
#pragma unroll
    for (int k = 0; k < R; ++k) hist[k] = X[k];
    fdiff_inv<CT, R>(hist);
    scan_dir<CT, R, TS>(v, hist, a, causal, first && clamp);
    if (have && !last) {
        // the carry was known up front: one scan, its tail is the completed tail
        fdiff_fwd<CT, R>(hist);
        CT* dst = dm.inc + (size_t)sidx * R * TS + tid;
#pragma unroll
        for (int k = 0; k < R; ++k) dst[k * TS] = hist[k];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_u32(dm.status + sidx, st_inc);
    }
}

template <typename CT, int R, int TS>
__global__ void __launch_bounds__(TS, (TS == 128 ? 3 : 8))
lb_tile_kernel(const __grid_constant__ LBTileParams<CT, R> p, const __grid_constant__ CUtensorMap tm_in,
               const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int NBOX = TS / 32;
    constexpr int BOX_BYTES = TS * 128;
    extern __shared__ __align__(16) unsigned char lbsmem_raw[];
    unsigned char* tile = lbsmem_raw + ((1024u - (smem_u32(lbsmem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + NBOX * BOX_BYTES);
    volatile uint32_t* sflag = reinterpret_cast<volatile uint32_t*>(bar + 1);     // [0] ticket, [1] predecessor ready

    const int tid = threadIdx.x;
    if (tid == 0) {
        sflag[0] = atomicInc(p.ticket, gridDim.x - 1);      // tiles in scan order; the last ticket resets the counter
        mbar_init(bar, 1);
    }
    __syncthreads();
    const uint32_t t = sflag[0];
    pdl_launch_dependents();
    // scan-order coordinates (x fastest) and the tile they denote in memory
    const int bxs = (int)(t % (uint32_t)p.nbx);
    const uint32_t rest = t / (uint32_t)p.nbx;
    const int bds = (int)(rest % (uint32_t)p.nbd);
    const int64_t o = rest / (uint32_t)p.nbd;
    const int bx = (p.x.nscan && !p.x.causal) ? p.nbx - 1 - bxs : bxs;
    const int bd = (p.d.nscan && !p.d.causal) ? p.nbd - 1 - bds : bds;
    const int x0 = bx * TS;
    const int y0 = (int)(o * p.Nd + (int64_t)bd * TS);

    pdl_wait();                                  // the input may come from the previous kernel; nothing is published before
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, x0 + bb * 32, y0, bar);
    }

    CT v[TS];
    const uint32_t cbase = smem_u32(tile) + (tid >> 5) * BOX_BYTES + (((tid & 31) >> 2) << 4) + ((tid & 3) << 2);
    const uint32_t rbase = smem_u32(tile) + tid * 128;
    const uint32_t rx = (tid & 7) << 4;
    auto load_col = [&]() {
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"((cbase ^ ((i & 7) << 4)) + i * 128));
            v[i] = *reinterpret_cast<CT*>(&w);
        }
    };
    auto load_row = [&]() {
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
    };

    mbar_wait(bar, 0);
    if (p.d.nscan) {
        // ---- column phase: thread tid owns column tid; predecessors are the tiles above (scan order) ----
        load_col();
        lb_phase<CT, R, TS>(v, p.d, bds, p.nbd, t, (uint32_t)p.nbx, p.epoch, p.clamp != 0, p.err, sflag, tid, load_col);
        const CT g = p.x.nscan ? (CT)1 : p.gain;
#pragma unroll
        for (int i = 0; i < TS; ++i) {
            const CT w = p.x.nscan ? v[i] : v[i] * g;
            asm volatile("st.shared.b32 [%0], %1;" :: "r"((cbase ^ ((i & 7) << 4)) + i * 128),
                         "r"(*reinterpret_cast<const uint32_t*>(&w)) : "memory");
        }
    }
    if (p.x.nscan) {
        // ---- row phase on the d-complete tile: thread tid owns row tid; predecessors are the tiles before it ----
        if (p.d.nscan) __syncthreads();
        load_row();
        lb_phase<CT, R, TS>(v, p.x, bxs, p.nbx, t, 1u, p.epoch, p.clamp != 0, p.err, sflag, tid, load_row);
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
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_store_2d(&tm_out, x0 + bb * 32, y0, tile + bb * BOX_BYTES);
        tma_store_commit_and_wait_read();
    }
}

// ---------------------------------------------------------------------------------------------
// long 1-D signals, single pass
// ---------------------------------------------------------------------------------------------
// y = M x with a per-lane matrix from a [R*R][32] table
template <typename CT, int R>
__device__ __forceinline__ void lb_lane_matvec(CT (&y)[R], const CT* table, const int lane, const CT (&x)[R])
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        CT acc = (CT)0;
#pragma unroll
        for (int kk = 0; kk < R; ++kk) acc = fmadd(__ldg(table + (k * R + kk) * 32 + lane), x[kk], acc);
        y[k] = acc;
    }
}
template <typename CT>
__device__ __forceinline__ CT lb_shfl_up(CT x, int d)
{
    const uint32_t r = __shfl_up_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&x), d);
    return *reinterpret_cast<const CT*>(&r);
}
template <typename CT>
__device__ __forceinline__ CT lb_shfl_xor(CT x, int d)
{
    const uint32_t r = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&x), d);
    return *reinterpret_cast<const CT*>(&r);
}

template <typename CT, int R>
__global__ void __launch_bounds__(128, 3)
lb_signal_kernel(const __grid_constant__ LBSignalParams<CT, R> p, const __grid_constant__ CUtensorMap tm_in,
                 const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int TS = 128, NBOX = 4, BOX_BYTES = TS * 128;
    extern __shared__ __align__(16) unsigned char lbsmem_raw[];
    unsigned char* tile = lbsmem_raw + ((1024u - (smem_u32(lbsmem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + NBOX * BOX_BYTES);
    volatile uint32_t* sflag = reinterpret_cast<volatile uint32_t*>(bar + 1);
    __shared__ CT swagg[4][R];        // inclusive tail of each warp's last row (zero carry into the warp)
    __shared__ CT sE0[4][R];          // carry entering each warp when nothing enters the tile
    __shared__ CT sX[R];              // carry entering the tile

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        sflag[0] = atomicInc(p.ticket, gridDim.x - 1);
        mbar_init(bar, 1);
    }
    __syncthreads();
    const uint32_t t = sflag[0];                                  // tile index in scan order
    pdl_launch_dependents();
    const bool causal = p.causal != 0;
    const uint32_t pos = t % (uint32_t)p.tiles_per_signal;        // tile of its signal, scan order
    const bool last_of_signal = pos == (uint32_t)p.tiles_per_signal - 1;
    const uint32_t mt = causal ? t : gridDim.x - 1 - t;           // the tile in memory
    const int y0 = (int)(mt * TS);
    // thread order = scan order: thread tid scans row `row`, after the row of thread tid - 1
    const int row = causal ? tid : TS - 1 - tid;
    const bool closed = pos == 0 && tid == 0;                     // the scan starts at the signal's border here
    const uint32_t st_agg = (p.epoch << 2) | LB_AGGREGATE, st_inc = (p.epoch << 2) | LB_INCLUSIVE;

    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, bb * 32, y0, bar);
    }
    CT a[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) a[k] = p.a[k];
    CT v[TS];
    const uint32_t rbase = smem_u32(tile) + row * 128;
    const uint32_t rx = (row & 7) << 4;
    auto load_row = [&]() {
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
    };
    mbar_wait(bar, 0);
    load_row();

    // ---- the row's own tail, then an inclusive scan over the rows of the warp (Kogge-Stone) ----
    CT T[R];
#pragma unroll
    for (int k = 0; k < R; ++k) T[k] = (CT)0;
    scan_dir<CT, R, TS>(v, T, a, causal, closed && p.clamp);
    fdiff_fwd<CT, R>(T);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        CT y[R];
#pragma unroll
        for (int k = 0; k < R; ++k) y[k] = lb_shfl_up<CT>(T[k], 1 << i);
        if (lane >= (1 << i)) lb_matvec_acc<CT, R>(T, p.Pstep[i], y);
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < R; ++k) swagg[w][k] = T[k];
    }
    __syncthreads();

    if (w == 0) {
        // ---- carries entering the warps (zero carry into the tile), the tile's aggregate ----
        CT E[R];
#pragma unroll
        for (int k = 0; k < R; ++k) E[k] = (CT)0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < R; ++k) sE0[u][k] = E[k];
            }
            CT n[R];
#pragma unroll
            for (int k = 0; k < R; ++k) n[k] = swagg[u][k];
            lb_matvec_acc<CT, R>(n, p.Pwarp, E);
#pragma unroll
            for (int k = 0; k < R; ++k) E[k] = n[k];
        }
        // E is the aggregate of the tile
        CT X[R];
#pragma unroll
        for (int k = 0; k < R; ++k) X[k] = (CT)0;
        if (pos != 0) {
            if (!last_of_signal) {
                if (lane < R) {
                    CT mine = (CT)0;
#pragma unroll
                    for (int k = 0; k < R; ++k) if (lane == k) mine = E[k];
                    p.agg[(size_t)t * R + lane] = mine;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(p.status + t, st_agg);
            }
            // ---- look back: lane k examines the tile k+1 (+32, +64 ...) before this one ----
            for (uint32_t base = 1, win = 0; ; base += 32, ++win) {
                const uint32_t dist = base + (uint32_t)lane;
                const bool active = dist <= pos;
                uint32_t state = LB_NONE;
                if (active) state = lb_wait_status(p.status + (t - dist), p.epoch, p.err);
                const uint32_t incl = __ballot_sync(0xffffffffu, active && state == LB_INCLUSIVE);
                const int firsti = incl ? __ffs(incl) - 1 : 32;                  // nearest complete predecessor of the window
                CT c[R];
#pragma unroll
                for (int k = 0; k < R; ++k) c[k] = (CT)0;
                if (active && lane <= firsti) {
                    const CT* src = (state == LB_INCLUSIVE ? p.inc : p.agg) + (size_t)(t - dist) * R;
                    CT y[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) y[k] = ld_cg(src + k);
                    lb_lane_matvec<CT, R>(c, p.Qpow, lane, y);                  // Q^lane * (aggregate | inclusive)
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int k = 0; k < R; ++k) c[k] = c[k] + lb_shfl_xor<CT>(c[k], off);
                for (uint32_t m = 0; m < win; ++m) {                            // the window lies 32 * win tiles back
                    CT n[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) n[k] = (CT)0;
                    lb_matvec_acc<CT, R>(n, p.Q32, c);
#pragma unroll
                    for (int k = 0; k < R; ++k) c[k] = n[k];
                }
#pragma unroll
                for (int k = 0; k < R; ++k) X[k] = X[k] + c[k];
                if (incl || base + 31 >= pos) break;
            }
        }
        if (!last_of_signal) {
            lb_matvec_acc<CT, R>(E, p.Q, X);                                     // completed tail of the tile
            if (lane < R) {
                CT mine = (CT)0;
#pragma unroll
                for (int k = 0; k < R; ++k) if (lane == k) mine = E[k];
                p.inc[(size_t)t * R + lane] = mine;
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(p.status + t, st_inc);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < R; ++k) sX[k] = X[k];
        }
    }
    __syncthreads();

    // ---- the carry entering this row: P^lane * (carry entering the warp) + inclusive tail of the row before ----
    CT Ew[R];
#pragma unroll
    for (int k = 0; k < R; ++k) Ew[k] = sX[k];
    for (int u = 0; u < w; ++u) {
        CT n[R];
#pragma unroll
        for (int k = 0; k < R; ++k) n[k] = (CT)0;
        lb_matvec_acc<CT, R>(n, p.Pwarp, Ew);
#pragma unroll
        for (int k = 0; k < R; ++k) Ew[k] = n[k];
    }
#pragma unroll
    for (int k = 0; k < R; ++k) Ew[k] = Ew[k] + sE0[w][k];
    CT hist[R];
    lb_lane_matvec<CT, R>(hist, p.Plane, lane, Ew);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const CT prev = lb_shfl_up<CT>(T[k], 1);
        if (lane > 0) hist[k] = hist[k] + prev;
    }
    fdiff_inv<CT, R>(hist);

    load_row();
    scan_dir<CT, R, TS>(v, hist, a, causal, closed && p.clamp);
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
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_store_2d(&tm_out, bb * 32, y0, tile + bb * BOX_BYTES);
        tma_store_commit_and_wait_read();
    }
}

} // namespace rfb
