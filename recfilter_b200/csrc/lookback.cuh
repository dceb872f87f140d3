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
 *      already running or finished: no deadlock whatever the block scheduler does; the ticket also carries the
 *      number of the launch, see lb_take_ticket),
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

/*
 * Publication protocol: a carry vector travels as 16-byte chunks {3 values, tag}, tag = (epoch << 2) | state.
 * An aligned 16-byte store / load is one transaction, so every chunk validates itself: no status word, no
 * fence, no barrier between the payload and its flag.  A vector of R values takes (R + 2) / 3 chunks; a reader
 * accepts a vector when all its chunks carry the same tag of this launch.  A record is ONE such vector: the
 * owner first stores its aggregate there and later overwrites it with its inclusive vector (a reader that
 * catches the overwrite half way sees different tags and polls again).  The epoch changes with every launch,
 * so the records are never cleared: a chunk of an older launch reads as "nothing published".
 */
template <int R> struct LBChunks { static constexpr int N = (R + 2) / 3; };

__device__ __forceinline__ uint4 lb_ld_chunk(const uint4* p)
{
    uint4 r;                                  // volatile: served by L2, never by a stale L1 line
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void lb_st_chunk(uint4* p, uint4 v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <typename CT> __device__ __forceinline__ uint32_t lb_bits(CT x) { return *reinterpret_cast<uint32_t*>(&x); }
template <typename CT> __device__ __forceinline__ CT lb_val(uint32_t x) { return *reinterpret_cast<CT*>(&x); }

// write the vector h as chunks with the given tag
template <typename CT, int R>
__device__ __forceinline__ void lb_publish(uint4* rec, const CT (&h)[R], const uint32_t tag)
{
#pragma unroll
    for (int c = 0; c < LBChunks<R>::N; ++c) {
        uint4 q;
        q.x = lb_bits<CT>(h[3 * c]);
        q.y = 3 * c + 1 < R ? lb_bits<CT>(h[(3 * c + 1) % R]) : 0u;
        q.z = 3 * c + 2 < R ? lb_bits<CT>(h[(3 * c + 2) % R]) : 0u;
        q.w = tag;
        lb_st_chunk(rec + c, q);
    }
}
template <int R>
__device__ __forceinline__ void lb_load_record(const uint4* rec, uint4 (&q)[LBChunks<R>::N])
{
#pragma unroll
    for (int c = 0; c < LBChunks<R>::N; ++c) q[c] = lb_ld_chunk(rec + c);
}
// state of a loaded record (LB_NONE when it is not a complete vector of this launch); y gets the vector
template <typename CT, int R>
__device__ __forceinline__ uint32_t lb_decode(const uint4 (&q)[LBChunks<R>::N], CT (&y)[R], const uint32_t epoch)
{
    const uint32_t tag = q[0].w;
    bool ok = (tag >> 2) == epoch && (tag & 3u) != LB_NONE;
#pragma unroll
    for (int c = 0; c < LBChunks<R>::N; ++c) {
        ok = ok && q[c].w == tag;
        y[3 * c] = lb_val<CT>(q[c].x);
        if (3 * c + 1 < R) y[(3 * c + 1) % R] = lb_val<CT>(q[c].y);
        if (3 * c + 2 < R) y[(3 * c + 2) % R] = lb_val<CT>(q[c].z);
    }
    return ok ? (tag & 3u) : (uint32_t)LB_NONE;
}
/*
 * Wait until the record holds an aggregate or an inclusive vector of this launch; returns the state with the
 * vector in y.  Never hangs the device: gives up after LB_SPIN_LIMIT polls (and at once when another thread
 * already has), raising the error flag that rf_plan_check reports.
 */
template <typename CT, int R>
__device__ __forceinline__ uint32_t lb_wait_record(const uint4* rec, CT (&y)[R], const uint32_t epoch, uint32_t* err)
{
    uint32_t spins = 0;
    while (true) {
        uint4 q[LBChunks<R>::N];
        lb_load_record<R>(rec, q);
        const uint32_t state = lb_decode<CT, R>(q, y, epoch);
        if (state != LB_NONE) return state;
        if (++spins > LB_SPIN_LIMIT || ((spins & 1023u) == 0u && *reinterpret_cast<volatile uint32_t*>(err) != 0u)) {
            *err = 1u;
#pragma unroll
            for (int k = 0; k < R; ++k) y[k] = (CT)0;
            return LB_INCLUSIVE;
        }
        if (spins > 4) __nanosleep(32);
    }
}

/*
 * Tickets come from a 64-bit counter in device memory that is never reset.  Every launch uses a ticket space of
 * P = 2^shift >= tiles tickets: ticket >> shift is the number of the launch (the epoch of the record tags), ticket & (P-1)
 * the tile; the CTA that draws the last tile skips the rest of the space.  One atomic gives both numbers without a
 * division, and nothing about a launch is baked into its parameters, so a captured CUDA graph can be replayed (every
 * replay is a new epoch).  The tickets of one launch are contiguous: a dependent launch only starts once every CTA of the
 * launch before has taken its ticket (tickets are taken before griddepcontrol.launch_dependents).
 */
__device__ __forceinline__ void lb_take_ticket(unsigned long long* ctr, const uint32_t tiles, volatile uint32_t* sflag)
{
    const int shift = 32 - __clz(tiles - 1u);                      // P = 2^shift >= tiles (tiles >= 1; tiles == 1: shift 0)
    const unsigned long long raw = atomicAdd(ctr, 1ull);
    const uint32_t t = (uint32_t)(raw & ((1ull << shift) - 1ull));
    if (t == tiles - 1u && (1ull << shift) != (unsigned long long)tiles) atomicAdd(ctr, (1ull << shift) - (unsigned long long)tiles);
    sflag[0] = t;
    sflag[2] = (uint32_t)((raw >> shift) % 0x3fffffffull) + 1u;    // never 0: cleared records carry tag 0
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
// x <- M x
template <typename CT, int R>
__device__ __forceinline__ void lb_matvec_inplace(CT (&x)[R], const CT* m)
{
    CT n[R];
#pragma unroll
    for (int k = 0; k < R; ++k) n[k] = (CT)0;
    lb_matvec_acc<CT, R>(n, m, x);
#pragma unroll
    for (int k = 0; k < R; ++k) x[k] = n[k];
}

// ---------------------------------------------------------------------------------------------
// 2-D single pass
// ---------------------------------------------------------------------------------------------
/*
 * One dimension of a tile: on entry every thread holds its line of the tile in v (the column of thread tid
 * in the d phase, the row in the x phase); on return v is the line filtered with the carry that enters the
 * tile.  `reload` re-reads the line from shared memory (the unfiltered values: needed when the line was first
 * scanned with zero history for the aggregate).  Every line has its own record per tile (rec[tile][line]), so
 * a thread depends on nobody but the thread that owns the same line in the tiles before: no barriers in here.
 *   sidx    index of this tile in scan-order coordinates, pstride the index distance to the previous tile of the line
 *   j       position of the tile along the dimension in scan order (0 = starts at the closed border)
 */
template <typename CT, int R, int TS, typename Reload>
__device__ __forceinline__ void lb_phase(CT (&v)[TS], const LBDim<CT, R>& dm, const int j, const int nb, const uint32_t sidx,
                                         const uint32_t pstride, const uint32_t epoch, const bool clamp, uint32_t* err,
                                         const int tid, Reload reload, const uint4 (&peek)[LBChunks<R>::N])
{
    constexpr int NCH = LBChunks<R>::N;
    constexpr int B = 8 / NCH;                           // predecessors fetched per round trip
    const bool causal = dm.causal != 0;
    const bool first = j == 0, last = j == nb - 1;
    const uint32_t tag_agg = (epoch << 2) | LB_AGGREGATE, tag_inc = (epoch << 2) | LB_INCLUSIVE;
    uint4* const recs = reinterpret_cast<uint4*>(dm.rec);
    auto rec_of = [&](int q) -> uint4* { return recs + ((size_t)(sidx - (uint32_t)q * pstride) * TS + tid) * NCH; };
    uint4* const mine = rec_of(0);
    CT a[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) a[k] = dm.a[k];
    CT X[R];                                             // carry entering the tile, difference basis
#pragma unroll
    for (int k = 0; k < R; ++k) X[k] = (CT)0;

    bool have = first;
    if (!first) {
        // the tile before this one may already be complete (`peek`: its record, fetched by the caller behind other
        // work); decided per warp, so that a warp scans once or twice as a whole
        CT y[R];
        const bool ok = lb_decode<CT, R>(peek, y, epoch) == LB_INCLUSIVE;
        have = __all_sync(0xffffffffu, ok);
        if (have) {
#pragma unroll
            for (int k = 0; k < R; ++k) X[k] = y[k];
        }
    }
    if (!have) {
        // aggregate: the line's own tail (zero history), published before anything is waited for
        CT h[R];
#pragma unroll
        for (int k = 0; k < R; ++k) h[k] = (CT)0;
        scan_dir<CT, R, TS>(v, h, a, causal, false);
        fdiff_fwd<CT, R>(h);
        if (!last) lb_publish<CT, R>(mine, h, tag_agg);
        // look back: X = sum over the tiles before of P^(q-1) * (aggregate | inclusive), ended by the nearest
        // inclusive; the records of B predecessors are fetched per round trip
        bool done = false;
        for (int q0 = 1; q0 <= j && !done; q0 += B) {
            uint4 ch[B][NCH];
#pragma unroll
            for (int i = 0; i < B; ++i)
                if (q0 + i <= j) lb_load_record<R>(rec_of(q0 + i), ch[i]);
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int q = q0 + i;
                if (!done && q <= j) {
                    CT y[R];
                    uint32_t state = lb_decode<CT, R>(ch[i], y, epoch);
                    if (state == LB_NONE) state = lb_wait_record<CT, R>(rec_of(q), y, epoch, err);
                    lb_matvec_acc<CT, R>(X, dm.Ppow + (size_t)(q - 1) * R * R, y);
                    done = state == LB_INCLUSIVE;
                }
            }
        }
        if (!last) {
            // completed tail = aggregate + P * carry: successors need not wait for the re-scan
            lb_matvec_acc<CT, R>(h, dm.P, X);
            lb_publish<CT, R>(mine, h, tag_inc);
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
        lb_publish<CT, R>(mine, hist, tag_inc);
    }
}

// ticket -> tile: tiles of an image are handed out along anti-diagonals of the scan-order grid (diagonal s lists
// its tiles by ascending bds), so the tiles a tile waits for (left, above) are a whole diagonal older.  Closed
// form (no table, no memory latency on the critical path of a CTA): growing diagonals, full-length diagonals,
// shrinking diagonals.
__device__ __forceinline__ uint32_t lb_tri_inv(const uint32_t q)          // largest s with s (s + 1) / 2 <= q
{
    uint32_t s = (uint32_t)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while ((s + 1) * (s + 2) / 2 <= q) ++s;
    while (s * (s + 1) / 2 > q) --s;
    return s;
}
__device__ __forceinline__ void lb_decode_ticket(const uint32_t t, const int nbx, const int nbd, const bool rows_first,
                                                 int& bxs, int& bds, int64_t& o)
{
    const uint32_t tpi = (uint32_t)nbx * (uint32_t)nbd;
    const uint32_t oi = t / tpi, r = t - oi * tpi;
    o = oi;
    if (rows_first) { bds = (int)(r / (uint32_t)nbx); bxs = (int)(r - (uint32_t)bds * (uint32_t)nbx); return; }
    const uint32_t m = (uint32_t)min(nbx, nbd), M = (uint32_t)max(nbx, nbd);
    const uint32_t tri = m * (m - 1) / 2, mid = (M - m + 1) * m;
    uint32_t sd, i;
    if (r < tri) { sd = lb_tri_inv(r); i = r - sd * (sd + 1) / 2; }
    else if (r < tri + mid) { const uint32_t q = r - tri; sd = m - 1 + q / m; i = q % m; }
    else { const uint32_t q = tpi - 1 - r, s2 = lb_tri_inv(q); sd = (uint32_t)(nbx + nbd - 2) - s2; i = s2 - (q - s2 * (s2 + 1) / 2); }
    bds = max(0, (int)sd - (nbx - 1)) + (int)i;
    bxs = (int)sd - bds;
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
    volatile uint32_t* sflag = reinterpret_cast<volatile uint32_t*>(bar + 1);     // [0] ticket, [2] epoch of this launch

    const int tid = threadIdx.x;
    if (tid == 0) {
        lb_take_ticket(p.ticket, gridDim.x, sflag);
        mbar_init(bar, 1);
    }
    __syncthreads();
    const uint32_t t = sflag[0];
    const uint32_t epoch = sflag[2];
    pdl_launch_dependents();
    // scan-order coordinates and the tile they denote in memory
    int bxs, bds; int64_t o;
    lb_decode_ticket(t, p.nbx, p.nbd, p.rows_first != 0, bxs, bds, o);
    const int bx = (p.x.nscan && !p.x.causal) ? p.nbx - 1 - bxs : bxs;
    const int bd = (p.d.nscan && !p.d.causal) ? p.nbd - 1 - bds : bds;
    const int x0 = bx * TS;
    const int y0 = (int)(o * p.Nd + (int64_t)bd * TS);
    const uint32_t sidx = ((uint32_t)o * (uint32_t)p.nbd + (uint32_t)bds) * (uint32_t)p.nbx + (uint32_t)bxs;

    pdl_wait();                                  // the input may come from the previous kernel; nothing is published before
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, x0 + bb * 32, y0, bar);
        // optional: the tile that will be handed out `prefetch` tickets from now starts its way into L2
        const uint32_t t2 = t + (uint32_t)p.prefetch;
        if (p.prefetch > 0 && t2 < gridDim.x) {
            int bxs2, bds2; int64_t o2;
            lb_decode_ticket(t2, p.nbx, p.nbd, p.rows_first != 0, bxs2, bds2, o2);
            const int bx2 = (p.x.nscan && !p.x.causal) ? p.nbx - 1 - bxs2 : bxs2;
            const int bd2 = (p.d.nscan && !p.d.causal) ? p.nbd - 1 - bds2 : bds2;
#pragma unroll
            for (int bb = 0; bb < NBOX; ++bb) tma_prefetch_2d(&tm_in, bx2 * TS + bb * 32, (int)(o2 * p.Nd + (int64_t)bd2 * TS));
        }
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

    // the record of the tile before this one along d travels while the tile itself is still on its way
    constexpr int NCH = LBChunks<R>::N;
    uint4 peek[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) peek[c] = make_uint4(0u, 0u, 0u, 0u);
    if (p.d.nscan && bds > 0)
        lb_load_record<R>(reinterpret_cast<const uint4*>(p.d.rec) + ((size_t)(sidx - (uint32_t)p.nbx) * TS + tid) * NCH, peek);
    mbar_wait(bar, 0);
    if (p.d.nscan) {
        // ---- column phase: thread tid owns column tid; predecessors are the tiles above (scan order) ----
        load_col();
        lb_phase<CT, R, TS>(v, p.d, bds, p.nbd, sidx, (uint32_t)p.nbx, epoch, p.clamp != 0, p.err, tid, load_col, peek);
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
        if (bxs > 0)                                      // (fetched behind the barrier and the row loads)
            lb_load_record<R>(reinterpret_cast<const uint4*>(p.x.rec) + ((size_t)(sidx - 1u) * TS + tid) * NCH, peek);
        if (p.d.nscan) __syncthreads();
        load_row();
        lb_phase<CT, R, TS>(v, p.x, bxs, p.nbx, sidx, 1u, epoch, p.clamp != 0, p.err, tid, load_row, peek);
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

/*
 * One CTA = 32 * NW consecutive rows of 128 samples of one signal, thread = row, thread order = scan order
 * (NW = warps per CTA: 4, 2 or 1 -- smaller tiles mean more independent CTAs per SM to hide the load and
 * look-back latencies behind, at the same work per thread).  A row is scanned in four chunks of 32 samples
 * straight from / to shared memory (small code, few registers): pass 0 with zero history for the row's tail,
 * pass 1 from the carry, storing.  Between the passes: Kogge-Stone over the rows of a warp, warps chained
 * through shared memory, then the look-back over the tiles before -- by all warps at once, warp w examining
 * the tiles 32w+1 .. 32w+32 back (lane = tile), so one round trip covers 32 * NW predecessors.
 */
template <typename CT, int R, int NW>
__global__ void __launch_bounds__(32 * NW, 12 / NW)
lb_signal_kernel(const __grid_constant__ LBSignalParams<CT, R> p, const __grid_constant__ CUtensorMap tm_in,
                 const __grid_constant__ CUtensorMap tm_out)
{
    constexpr int ROWS = 32 * NW, NBOX = 4, BOX_BYTES = ROWS * 128;
    extern __shared__ __align__(16) unsigned char lbsmem_raw[];
    unsigned char* tile = lbsmem_raw + ((1024u - (smem_u32(lbsmem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + NBOX * BOX_BYTES);
    volatile uint32_t* sflag = reinterpret_cast<volatile uint32_t*>(bar + 1);
    __shared__ CT swagg[NW][R];       // inclusive tail of each warp's last row (zero carry into the warp)
    __shared__ CT sS[NW][R];          // look-back: partial sum of each warp's window
    __shared__ uint32_t sI[NW];       //            ... and whether the window held an inclusive vector

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        lb_take_ticket(p.ticket, gridDim.x, sflag);
        mbar_init(bar, 1);
    }
    __syncthreads();
    const uint32_t t = sflag[0];                                  // tile index in scan order
    const uint32_t epoch = sflag[2];
    pdl_launch_dependents();
    const bool causal = p.causal != 0;
    const uint32_t pos = t % (uint32_t)p.tiles_per_signal;        // tile of its signal, scan order
    const bool last_of_signal = pos == (uint32_t)p.tiles_per_signal - 1;
    const uint32_t mt = causal ? t : gridDim.x - 1 - t;           // the tile in memory
    const int y0 = (int)(mt * ROWS);
    // thread order = scan order: thread tid scans row `row`, after the row of thread tid - 1
    const int row = causal ? tid : ROWS - 1 - tid;
    const bool closed = pos == 0 && tid == 0;                     // the scan starts at the signal's border here
    uint4* const recs = reinterpret_cast<uint4*>(p.rec);
    const uint32_t tag_agg = (epoch << 2) | LB_AGGREGATE, tag_inc = (epoch << 2) | LB_INCLUSIVE;

    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, bb * 32, y0, bar);
        const uint32_t t2 = t + (uint32_t)p.prefetch;          // optional L2 prefetch of a tile handed out later
        if (p.prefetch > 0 && t2 < gridDim.x) {
            const uint32_t mt2 = causal ? t2 : gridDim.x - 1 - t2;
#pragma unroll
            for (int bb = 0; bb < NBOX; ++bb) tma_prefetch_2d(&tm_in, bb * 32, (int)(mt2 * ROWS));
        }
    }
    CT a[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) a[k] = p.a[k];
    const uint32_t rbase = smem_u32(tile) + row * 128;
    const uint32_t rx = (row & 7) << 4;
    mbar_wait(bar, 0);

    // scan the row from / to shared memory, 32 samples (one box) at a time; h: history in, tail out
    auto scan_row = [&](CT (&h)[R], auto store, const int first_chunk) {
#pragma unroll 1
        for (int cc = first_chunk; cc < NBOX; ++cc) {
            const int c = causal ? cc : NBOX - 1 - cc;
            CT v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 q;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                             : "r"(rbase + c * BOX_BYTES + ((i << 4) ^ rx)));
                v[i * 4 + 0] = *reinterpret_cast<const CT*>(&q.x);
                v[i * 4 + 1] = *reinterpret_cast<const CT*>(&q.y);
                v[i * 4 + 2] = *reinterpret_cast<const CT*>(&q.z);
                v[i * 4 + 3] = *reinterpret_cast<const CT*>(&q.w);
            }
            scan_dir<CT, R, 32>(v, h, a, causal, closed && p.clamp && cc == 0);
            if constexpr (decltype(store)::value) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 q;
                    *reinterpret_cast<CT*>(&q.x) = v[i * 4 + 0] * p.gain;
                    *reinterpret_cast<CT*>(&q.y) = v[i * 4 + 1] * p.gain;
                    *reinterpret_cast<CT*>(&q.z) = v[i * 4 + 2] * p.gain;
                    *reinterpret_cast<CT*>(&q.w) = v[i * 4 + 3] * p.gain;
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(rbase + c * BOX_BYTES + ((i << 4) ^ rx)), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
                }
            }
        }
    };

    CT h[R];
#pragma unroll
    for (int k = 0; k < R; ++k) h[k] = (CT)0;
    // pass 0: the row's own tail.  Short-memory filters: the tail does not see the first chunks of the row (their
    // weight in it is below 1e-12, decided by the planner from the fp64 impulse response), so they are not scanned
    scan_row(h, std::false_type(), p.pass0_first_chunk);

    // ---- inclusive scan over the rows of the warp (Kogge-Stone), difference basis ----
    CT T[R];
#pragma unroll
    for (int k = 0; k < R; ++k) T[k] = h[k];
    fdiff_fwd<CT, R>(T);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        CT y[R];
#pragma unroll
        for (int k = 0; k < R; ++k) y[k] = lb_shfl_up<CT>(T[k], 1 << i);
        if (lane >= (1 << i)) lb_matvec_acc<CT, R>(T, p.Pstep[i], y);
    }
    // ---- carry entering this thread's warp when nothing enters the tile (Ew), the tile's aggregate (E) ----
    CT E[R], Ew[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { E[k] = (CT)0; Ew[k] = (CT)0; }
    if constexpr (NW > 1) {
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < R; ++k) swagg[w][k] = T[k];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < NW; ++u) {
            if (u == w) {
#pragma unroll
                for (int k = 0; k < R; ++k) Ew[k] = E[k];
            }
            if (u <= w || w == 0) {                                // warp w needs E_w; warp 0 also the tile's aggregate
                CT n[R];
#pragma unroll
                for (int k = 0; k < R; ++k) n[k] = swagg[u][k];
                lb_matvec_acc<CT, R>(n, p.Pwarp, E);
#pragma unroll
                for (int k = 0; k < R; ++k) E[k] = n[k];
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < R; ++k) E[k] = lb_val<CT>(__shfl_sync(0xffffffffu, lb_bits<CT>(T[k]), 31));   // one warp: its last row's tail
    }
    // ---- look back ----
    CT X[R];
#pragma unroll
    for (int k = 0; k < R; ++k) X[k] = (CT)0;
    if (p.depth1) {
        // short memory across a whole tile (|Q| < 1e-12: any stable filter whose poles are not within ~1e-3 of the unit
        // circle): the carry entering the tile is the aggregate of the tile before it, nothing further back counts
        // -- one record to wait for, no inclusive vectors
        if (!last_of_signal && tid == 0) lb_publish<CT, R>(recs + (size_t)t * LB_SIGNAL_REC_CHUNKS, E, tag_agg);
        if (pos != 0) lb_wait_record<CT, R>(recs + (size_t)(t - 1) * LB_SIGNAL_REC_CHUNKS, X, epoch, p.err);
    } else if (pos != 0) {
        if (!last_of_signal && tid == 0) lb_publish<CT, R>(recs + (size_t)t * LB_SIGNAL_REC_CHUNKS, E, tag_agg);
        for (uint32_t round = 0; ; ++round) {
            const uint32_t dist = round * (uint32_t)ROWS + (uint32_t)tid + 1u;
            const bool active = dist <= pos;
            uint32_t state = LB_NONE;
            CT y[R];
#pragma unroll
            for (int k = 0; k < R; ++k) y[k] = (CT)0;
            if (active) state = lb_wait_record<CT, R>(recs + (size_t)(t - dist) * LB_SIGNAL_REC_CHUNKS, y, epoch, p.err);
            const uint32_t incl = __ballot_sync(0xffffffffu, active && state == LB_INCLUSIVE);
            const int firsti = incl ? __ffs(incl) - 1 : 32;                  // nearest complete predecessor of the window
            CT c[R];
#pragma unroll
            for (int k = 0; k < R; ++k) c[k] = (CT)0;
            if (active && lane <= firsti) lb_lane_matvec<CT, R>(c, p.Qpow, lane, y);   // Q^lane * (aggregate | inclusive)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int k = 0; k < R; ++k) c[k] = c[k] + lb_shfl_xor<CT>(c[k], off);
            CT Y[R];
            bool any_incl;
            if constexpr (NW > 1) {
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < R; ++k) sS[w][k] = c[k];
                    sI[w] = incl;
                }
                __syncthreads();
                // windows in order: Y = S_0 + Q32 (S_1 + Q32 (S_2 + ...)), cut after the first window with an inclusive
                int wl = NW - 1;
#pragma unroll
                for (int u = NW - 1; u >= 0; --u) if (sI[u]) wl = u;
#pragma unroll
                for (int k = 0; k < R; ++k) Y[k] = sS[wl][k];
                for (int u = wl - 1; u >= 0; --u) {
                    lb_matvec_inplace<CT, R>(Y, p.Q32);
#pragma unroll
                    for (int k = 0; k < R; ++k) Y[k] = Y[k] + sS[u][k];
                }
                any_incl = false;
#pragma unroll
                for (int u = 0; u < NW; ++u) any_incl = any_incl || sI[u] != 0u;
                __syncthreads();                                              // sS / sI are reused by the next round
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) Y[k] = c[k];
                any_incl = incl != 0u;
            }
            for (uint32_t m = 0; m < (uint32_t)NW * round; ++m) lb_matvec_inplace<CT, R>(Y, p.Q32);   // the round lies 32 * NW * round tiles back
#pragma unroll
            for (int k = 0; k < R; ++k) X[k] = X[k] + Y[k];
            if (any_incl || round * (uint32_t)ROWS + (uint32_t)ROWS >= pos) break;
        }
    }
    if (!p.depth1 && !last_of_signal && tid == 0) {
        lb_matvec_acc<CT, R>(E, p.Q, X);                                     // completed tail of the tile
        lb_publish<CT, R>(recs + (size_t)t * LB_SIGNAL_REC_CHUNKS, E, tag_inc);
    }
    // ---- the carry entering this row: P^lane * (carry entering the warp) + inclusive tail of the row before ----
    for (int u = 0; u < w; ++u) lb_matvec_inplace<CT, R>(X, p.Pwarp);
#pragma unroll
    for (int k = 0; k < R; ++k) Ew[k] = Ew[k] + X[k];
    lb_lane_matvec<CT, R>(h, p.Plane, lane, Ew);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const CT prev = lb_shfl_up<CT>(T[k], 1);
        if (lane > 0) h[k] = h[k] + prev;
    }
    fdiff_inv<CT, R>(h);

    scan_row(h, std::true_type(), 0);                             // pass 1: from the carry, scaled, stored
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_store_2d(&tm_out, bb * 32, y0, tile + bb * BOX_BYTES);
        tma_store_commit_and_wait_read();
    }
}

} // namespace rfb
