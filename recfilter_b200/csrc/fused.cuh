#pragma once
/*
 * fused.cuh -- the fast path: fused two-dimension tile kernels with register-resident
 * lines and a segmented carry algebra in a difference basis (fp32 / the u32 ring; the matrices
 * are built in fp64 on the host).  Also the tile kernels' "signal mode" for long 1-D signals.
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
 *   fcrossA_kernel          A = L_x * CY  per tile (cross-dimension residual, part 1)
 *   fchain_kernel (x)       adds G_y * A to the x tails on the fly (part 2), then chains
 *   fused_tile_kernel<P2>   re-scan from the completed carries, scale, store
 */
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <type_traits>
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

/*
 * The same scan for a line of a partial tile (the last tile of a ragged extent): only the first `len`
 * samples are real, the rest is padding (TMA zero fill, or leftovers of an earlier scan) and is cleared
 * first.  A causal scan ends at sample len-1; an anticausal scan STARTS there, with the history it was
 * given (a carry from beyond, or the closed-border rule).  Taken by edge CTAs only.  (Inlined on purpose: a
 * call would force the register line of the common path into local memory.)
 */
template <typename CT, int R, int N, bool CAUSAL>
__device__ __forceinline__ void scan_line_partial(CT (&v)[N], CT (&h)[R], const CT (&a)[R + 1], const bool clampb, const int len)
{
#pragma unroll
    for (int i = 0; i < N; ++i) if (i >= len) v[i] = (CT)0;
#pragma unroll
    for (int p = 0; p < N; ++p) {
        const int i = CAUSAL ? p : N - 1 - p;              // compile-time: the line stays in registers
        if (i < len) {
            const bool first = CAUSAL ? (i == 0) : (i == len - 1);
            if (first && clampb) {
                const CT x0 = v[i] * a[0];
#pragma unroll
                for (int k = 0; k < R; ++k) h[k] = x0;
            }
            CT acc = v[i];
#pragma unroll
            for (int k = R; k >= 1; --k) acc = fmadd(a[k], h[k - 1], acc);
#pragma unroll
            for (int k = R - 1; k >= 1; --k) h[k] = h[k - 1];
            h[0] = acc;
            if (first && clampb) {
#pragma unroll
                for (int k = 1; k < R; ++k) h[k] = acc;
            }
            v[i] = acc;
        }
    }
}
template <typename CT, int R, int N, bool RAGGED>
__device__ __forceinline__ void scan_any(CT (&v)[N], CT (&h)[R], const CT (&a)[R + 1], const bool causal, const bool clampb,
                                         const int len)
{
    if (!RAGGED || len == N) {                               // every CTA but the edge ones
        if (causal) scan_line<CT, R, N, true >(v, h, a, clampb);
        else        scan_line<CT, R, N, false>(v, h, a, clampb);
    } else {
        if (causal) scan_line_partial<CT, R, N, true >(v, h, a, clampb, len);
        else        scan_line_partial<CT, R, N, false>(v, h, a, clampb, len);
    }
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): the kernels of a pass are launched back to back with
// cudaLaunchAttributeProgrammaticStreamSerialization; a kernel lets its successor start launching
// right away (its CTAs fill the SMs as ours retire and run their prologue) and itself waits for
// the complete, flushed predecessor before it touches anything the predecessor wrote.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait()              { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// L2 prefetch of a box (no shared memory, no barrier): lets a CTA pull in the tile of the CTA that will follow it
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int x, int y)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(tmap), "r"(x), "r"(y) : "memory");
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
// 4-byte asynchronous global -> shared copy (LDGSTS): no register, no scoreboard stall at the issue point
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// tile position class along a dimension
__device__ __forceinline__ int ftile_variant(int j, int nb)
{
    if (nb == 1) return V_SINGLE;
    if (j == 0) return V_FIRST;
    if (j == nb - 1) return V_LAST;
    return V_INTERIOR;
}

// y = D^-1 (M' (D c)): same-dimension residual of scan 0's carry in the tail of scan 1, M' = m[(var*S + 0)*S + 1]
template <typename CT, int R>
__device__ __forceinline__ void flocal_residual(const CT* m, const CT (&c)[R], CT (&y)[R])
{
    CT d[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { d[k] = c[k]; y[k] = (CT)0; }
#pragma unroll
    for (int mm = 1; mm < R; ++mm)
#pragma unroll
        for (int k = R - 1; k >= mm; --k) d[k] = d[k - 1] - d[k];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int kk = 0; kk < R; ++kk) y[k] = fmadd(__ldg(m + k * R + kk), d[kk], y[k]);
#pragma unroll
    for (int mm = R - 1; mm >= 1; --mm)
#pragma unroll
        for (int k = mm; k < R; ++k) y[k] = y[k - 1] - y[k];
}

// the same with M' in kernel-constant memory (no global load in front of the FMAs)
template <typename CT, int R>
__device__ __forceinline__ void flocal_residual_c(const CT (&m)[R * R], const CT (&c)[R], CT (&y)[R])
{
    CT d[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { d[k] = c[k]; y[k] = (CT)0; }
#pragma unroll
    for (int mm = 1; mm < R; ++mm)
#pragma unroll
        for (int k = R - 1; k >= mm; --k) d[k] = d[k - 1] - d[k];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int kk = 0; kk < R; ++kk) y[k] = fmadd(m[k * R + kk], d[kk], y[k]);
#pragma unroll
    for (int mm = R - 1; mm >= 1; --mm)
#pragma unroll
        for (int k = mm; k < R; ++k) y[k] = y[k - 1] - y[k];
}

// ---------------------------------------------------------------------------------------------
// P1 / P2: the tile kernel.
// Shared memory holds the tile as TS/32 TMA boxes of [TS rows][32 columns] (128 B rows) in the
// 128-byte swizzle: the 16-byte chunk c of row r lives at chunk c ^ (r & 7).  With that layout the
// thread-per-column accesses (32 lanes = the 32 words of one row) and the thread-per-row 128-bit
// accesses (8 lanes = 8 different chunks) are both bank-conflict free, without padding.
// ---------------------------------------------------------------------------------------------
// RAGGED: the pass has partial tiles (extents that are not multiples of TS); the instantiation without them is
// the exact full-tile kernel (the partial-tile code costs the common path ~15 % when it is compiled in)
// The body is shared by fused_tile_kernel (one launch per sweep: `b` is the block's tile, PDL on) and by
// fused_stream_kernel (STREAM: one launch does both sweeps, a CTA's work item comes from a ticket; `sy` tells a pass-2
// item which counters to wait for once its tile is on its way)
struct FStreamWait {
    const unsigned* cnt_p1;     // [rows] pass-1 tiles finished per tile row
    const unsigned* cnt_a;      // [rows] cross-residual blocks finished per tile row
    unsigned* err;
    int rows, row;              // tile rows of the whole stack, the row of this item
    unsigned need_p1, need_a;
};
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr unsigned FSTREAM_SPIN_LIMIT = 1u << 22;
// one thread: wait until tile rows [row - 2, row + 2] have finished pass 1 (and, need_a > 0, row's cross residuals exist)
__device__ __forceinline__ void fstream_wait(const FStreamWait& sy)
{
    unsigned spins = 0;
    bool ok = true;
    auto until = [&](const unsigned* c, unsigned need) {
        while (ok && ld_acquire_u32(c) < need) {
            __nanosleep(100);
            if (++spins > FSTREAM_SPIN_LIMIT) { ok = false; atomicExch(sy.err, 1u); }    // never hang: flag, go on
        }
    };
    for (int r = sy.row - 2; r <= sy.row + 2; ++r)
        if (r >= 0 && r < sy.rows) until(sy.cnt_p1 + r, sy.need_p1);
    if (sy.need_a) until(sy.cnt_a + sy.row, sy.need_a);
}

template <typename CT, int R, int TS, int MODE, bool RAGGED, bool STREAM, bool EPI = false>
__device__ __forceinline__ void fused_tile_body(const FusedParams<CT, R>& p, const CUtensorMap& tm_in, const CUtensorMap& tm_out,
                                                int64_t b, const FStreamWait& sy)
{
    constexpr int NBOX = TS / 32;
    constexpr int BOX_BYTES = TS * 128;
    extern __shared__ __align__(16) unsigned char fsmem_raw[];
    // the swizzle is a function of the shared-memory address: boxes must start on a 1024 B boundary
    unsigned char* tile = fsmem_raw + ((1024u - (smem_u32(fsmem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + NBOX * BOX_BYTES);
    CT* cbuf = reinterpret_cast<CT*>(tile + NBOX * BOX_BYTES + 16);      // P2: carries [d scans | x scans][R][TS]

    const int tid = threadIdx.x;
    const int bx = (int)(b % p.nbx); b /= p.nbx;
    const int bd = (int)(b % p.nbd);
    const int64_t o = b / p.nbd + p.o0;
    const int x0 = bx * TS;
    const int y0 = (int)(o * p.Nd + (int64_t)bd * TS);                  // row of the [No*Nd][Nx] matrix

    // partial tiles of a ragged extent: lenx / lend real samples per line, threads beyond them own padding lines
    const int lenx = RAGGED ? (int)min((int64_t)TS, p.Nx - x0) : TS;
    const int lend = (RAGGED && !p.signal) ? (int)min((int64_t)TS, p.Nd - (int64_t)bd * TS) : TS;
    const bool col_valid = !RAGGED || tid < lenx, row_valid = !RAGGED || tid < lend;

    if (!STREAM) pdl_launch_dependents();
    // does an x scan start at a closed border in the row of this thread?  (signal mode: a row continues the
    // previous row of the same signal, so only the first / last row of a signal is closed)
    const int64_t sig_row = p.signal ? ((int64_t)y0 + tid) % p.sig_rows : 0;
    auto x_closed = [&](bool causal) -> bool {
        if (p.signal) return causal ? (sig_row == 0) : (sig_row == p.sig_rows - 1);
        return causal ? (bx == 0 && p.x_lo_closed) : (bx == p.nbx - 1 && p.x_hi_closed);
    };
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (!STREAM) pdl_wait();                      // the input (and, for P2, the carries) may come from the previous kernel
    if (tid == 0) {
        mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
        for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, x0 + bb * 32, y0, bar);
        // the tile of the CTA that takes over this CTA's slot (`prefetch` blocks further on) starts its way into L2
        const int64_t nxt = (int64_t)blockIdx.x + p.prefetch;
        if (!STREAM && p.prefetch > 0 && nxt < (int64_t)gridDim.x) {
            int64_t b2 = p.reverse ? (int64_t)gridDim.x - 1 - nxt : nxt;
            const int bx2 = (int)(b2 % p.nbx); b2 /= p.nbx;
            const int bd2 = (int)(b2 % p.nbd);
            const int y2 = (int)((b2 / p.nbd + p.o0) * p.Nd + (int64_t)bd2 * TS);
#pragma unroll
            for (int bb = 0; bb < NBOX; ++bb) tma_prefetch_2d(&tm_in, bx2 * TS + bb * 32, y2);
        }
    }

    if (STREAM && MODE == FMODE_P2) {
        // the tile is on its way; the tails / cross residuals this item reads come from items with earlier tickets
        if (tid == 0) fstream_wait(sy);
        __syncthreads();
    }
    CT v[TS];
    CT hn[R];                                                          // history of the next scan

    if constexpr (MODE == FMODE_P2) {
      if (p.local) {
        // ---- short-memory pass 2: the carries are derived here from the tails of the neighbouring tiles ----
        // staging slots of R x TS words: [0, md + mx) the final carries (the layout the scan phases read), then one
        // temporary per dimension
        const int nfin = p.md + p.mx;
        CT* abuf = cbuf + (nfin + 2) * R * TS;                            // [3][mx][R][sdk]: A of the three x neighbours
        // row `tid` of the d response G (cross-dimension residual of the x tails): asked for first, used last
        CT grow[FMAX_SCANS * R];
        const bool crossx = p.A != nullptr && p.mx > 0 && p.md > 0;
        if (crossx) {
            const int vd = ftile_variant(bd, p.nbd);
#pragma unroll
            for (int n = 0; n < FMAX_SCANS * R; ++n)
                grow[n] = n < p.md * R ? __ldg(p.G + (((int64_t)vd * p.md + n / R) * TS + tid) * R + n % R) : (CT)0;
        }
        auto stage = [&](int slot, const CT* base, int64_t idx0, int64_t kstride, bool ok) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (ok) cp_async4(cbuf + (slot * R + k) * TS + tid, base + idx0 + k * kstride);
                else    cbuf[(slot * R + k) * TS + tid] = (CT)0;
            }
        };
        const int dd0 = (p.md > 0 && !p.sd.causal[0]) ? -1 : 1, dd1 = (p.md > 1 && !p.sd.causal[1]) ? -1 : 1;
        const int dx0 = (p.mx > 0 && !p.sx.causal[0]) ? -1 : 1, dx1 = (p.mx > 1 && !p.sx.causal[1]) ? -1 : 1;
        if (p.md > 0) {
            const int64_t ks = (int64_t)p.nbd * p.nly, ly = o * p.Nx + (int64_t)bx * TS + tid;
            auto src = [&](int slot, int s, int t) {
                stage(slot, p.TY, ((int64_t)s * R * p.nbd + t) * p.nly + ly, ks, col_valid && t >= 0 && t < p.nbd);
            };
            src(0, 0, bd - dd0);
            if (p.md > 1) { src(1, 1, bd - dd1); src(nfin, 0, bd - dd1 - dd0); }
        }
        if (p.mx > 0) {
            const int64_t ks = (int64_t)p.nbx * p.nlx, lx = o * p.Nd + (int64_t)bd * TS + tid;
            auto src = [&](int slot, int s, int t) {
                stage(slot, p.TX, ((int64_t)s * R * p.nbx + t) * p.nlx + lx, ks, row_valid && t >= 0 && t < p.nbx);
            };
            src(p.md + 0, 0, bx - dx0);
            if (p.mx > 1) { src(p.md + 1, 1, bx - dx1); src(nfin + 1, 0, bx - dx1 - dx0); }
            if (crossx) {
                // A of the x neighbours bx - dx0, bx - dx1, bx - dx1 - dx0 (16-byte copies by the first threads)
                const int per = p.mx * R * (p.sdk >> 2);                  // 16-byte chunks per tile
                for (int i = tid; i < 3 * per; i += TS) {
                    const int which = i / per, c = i - which * per;
                    const int t = which == 0 ? bx - dx0 : (which == 1 ? bx - dx1 : bx - dx1 - dx0);
                    if (t >= 0 && t < p.nbx && (which == 0 || p.mx > 1)) {
                        const CT* g = p.A + (((o * p.nbd + bd) * (int64_t)p.nbx + t) * p.mx) * R * p.sdk + (int64_t)c * 4;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(abuf + (which * per + c) * 4)), "l"(g) : "memory");
                    }
                }
            }
        }
        cp_async_wait_all();
        if (crossx) __syncthreads();                                     // the A matrices were staged by other threads
        if (p.md > 1 && col_valid) {
            const int t1 = bd - dd1;
            if (t1 >= 0 && t1 < p.nbd) {
                CT c0p[R], y[R];
#pragma unroll
                for (int k = 0; k < R; ++k) c0p[k] = cbuf[(nfin * R + k) * TS + tid];
                flocal_residual_c<CT, R>(p.Mld[ftile_variant(t1, p.nbd)], c0p, y);
#pragma unroll
                for (int k = 0; k < R; ++k) cbuf[(1 * R + k) * TS + tid] += y[k];
            }
        }
        if (p.mx > 0 && row_valid) {
            // T' = T + G_row * A[tile] for the three x sources (cross-dimension residual), then the combination along x
            auto cross = [&](int slot, int which, int s, int t) {
                if (!crossx || t < 0 || t >= p.nbx) return;
                const CT* ap = abuf + (which * p.mx + s) * R * p.sdk;
#pragma unroll
                for (int kx = 0; kx < R; ++kx) {
                    CT acc = cbuf[(slot * R + kx) * TS + tid];
#pragma unroll
                    for (int n = 0; n < FMAX_SCANS * R; ++n)
                        if (n < p.md * R) acc = fmadd(grow[n], ap[kx * p.sdk + n], acc);
                    cbuf[(slot * R + kx) * TS + tid] = acc;
                }
            };
            cross(p.md + 0, 0, 0, bx - dx0);
            if (p.mx > 1) {
                cross(p.md + 1, 1, 1, bx - dx1);
                cross(nfin + 1, 2, 0, bx - dx1 - dx0);
                const int t1 = bx - dx1;
                if (t1 >= 0 && t1 < p.nbx) {
                    CT c0p[R], y[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) c0p[k] = cbuf[((nfin + 1) * R + k) * TS + tid];
                    flocal_residual_c<CT, R>(p.Mlx[ftile_variant(t1, p.nbx)], c0p, y);
#pragma unroll
                    for (int k = 0; k < R; ++k) cbuf[((p.md + 1) * R + k) * TS + tid] += y[k];
                }
            }
        }
      } else {
        // every carry this thread will need (column tid, then row tid) starts its way into shared
        // memory now, behind the tile: no load latency is left inside the scan phases
        for (int s = 0; s < p.md; ++s) {
            const bool closed = p.sd.causal[s] ? (bd == 0 && p.d_lo_closed) : (bd == p.nbd - 1 && p.d_hi_closed);
            if (closed) continue;
            if (!col_valid) continue;
            const int64_t idx0 = ((int64_t)s * R * p.nbd + bd) * p.nly + o * p.Nx + (int64_t)bx * TS + tid;
#pragma unroll
            for (int k = 0; k < R; ++k) cp_async4(cbuf + (s * R + k) * TS + tid, p.CY + idx0 + k * (int64_t)p.nbd * p.nly);
        }
        for (int s = 0; s < p.mx; ++s) {
            const bool closed = x_closed(p.sx.causal[s] != 0);
            if (closed) continue;
            if (!row_valid) continue;
            const int64_t idx0 = ((int64_t)s * R * p.nbx + bx) * p.nlx + o * p.Nd + (int64_t)bd * TS + tid;
#pragma unroll
            for (int k = 0; k < R; ++k) cp_async4(cbuf + ((p.md + s) * R + k) * TS + tid, p.CX + idx0 + k * (int64_t)p.nbx * p.nlx);
        }
        cp_async_wait_all();           // (the wait retires behind the mbarrier wait of the tile in practice)
      }
    }

    if (p.md > 0) {
        // ---- column phase: thread tid owns column tid ----
        const int64_t ly = o * p.Nx + (int64_t)bx * TS + tid;
        const int64_t kstride = (int64_t)p.nbd * p.nly;
        auto load_cy = [&](int s) {
            const bool causal = p.sd.causal[s] != 0;
            const bool closed = causal ? (bd == 0 && p.d_lo_closed) : (bd == p.nbd - 1 && p.d_hi_closed);
#pragma unroll
            for (int k = 0; k < R; ++k) hn[k] = (MODE == FMODE_P2 && !closed && col_valid) ? cbuf[(s * R + k) * TS + tid] : (CT)0;
        };
        const uint32_t cbase = smem_u32(tile) + (tid >> 5) * BOX_BYTES + (((tid & 31) >> 2) << 4) + ((tid & 3) << 2);
        mbar_wait(bar, 0);
        load_cy(0);
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
            scan_any<CT, R, TS, RAGGED>(v, h, a, causal, closed && p.clamp, lend);
            if (MODE == FMODE_P1 && col_valid) {
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
            const bool closed = x_closed(p.sx.causal[s] != 0);
#pragma unroll
            for (int k = 0; k < R; ++k) hn[k] = (MODE == FMODE_P2 && !closed && row_valid) ? cbuf[((p.md + s) * R + k) * TS + tid] : (CT)0;
        };
        if (p.md > 0) __syncthreads(); else mbar_wait(bar, 0);
        load_cx(0);
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
        if (MODE == FMODE_P2 && EPI) {
            // every row is in registers, the tile buffer is free: the INPUT tile comes in again (TMA, an L2 hit) behind
            // the row scans, for the pointwise epilogue at the store
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(bar, NBOX * BOX_BYTES);
#pragma unroll
                for (int bb = 0; bb < NBOX; ++bb) tma_load_2d(tile + bb * BOX_BYTES, &tm_in, x0 + bb * 32, y0, bar);
            }
        }
        for (int s = 0; s < p.mx; ++s) {
            const bool causal = p.sx.causal[s] != 0;
            const bool closed = x_closed(causal);
            CT a[R + 1];
#pragma unroll
            for (int k = 0; k <= R; ++k) a[k] = p.sx.a[s][k];
            CT h[R];
#pragma unroll
            for (int k = 0; k < R; ++k) h[k] = hn[k];
            if (MODE == FMODE_P2 && s + 1 < p.mx) load_cx(s + 1);
            scan_any<CT, R, TS, RAGGED>(v, h, a, causal, closed && p.clamp, lenx);
            if (MODE == FMODE_P1 && row_valid) {
                const int64_t idx0 = ((int64_t)s * R * p.nbx + bx) * p.nlx + lx;
#pragma unroll
                for (int k = 0; k < R; ++k) p.TX[idx0 + k * kstride] = h[k];
            }
        }
        if constexpr (MODE == FMODE_P2) {
            if (EPI) {
                // out = epi_out * filtered + epi_in * input, the input row read back from the re-loaded tile
                mbar_wait(bar, 1);
                const CT go = p.gain * p.epi_out;
#pragma unroll
                for (int c4 = 0; c4 < TS / 4; ++c4) {
                    uint4 q;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                 : "r"(rbase + (c4 >> 3) * BOX_BYTES + ((((c4 & 7) << 4)) ^ rx)));
                    v[c4 * 4 + 0] = fmadd(v[c4 * 4 + 0], go, *reinterpret_cast<const CT*>(&q.x) * p.epi_in);
                    v[c4 * 4 + 1] = fmadd(v[c4 * 4 + 1], go, *reinterpret_cast<const CT*>(&q.y) * p.epi_in);
                    v[c4 * 4 + 2] = fmadd(v[c4 * 4 + 2], go, *reinterpret_cast<const CT*>(&q.z) * p.epi_in);
                    v[c4 * 4 + 3] = fmadd(v[c4 * 4 + 3], go, *reinterpret_cast<const CT*>(&q.w) * p.epi_in);
                }
            }
            const CT gs = EPI ? (CT)1 : p.gain;
            // scale, rows back to shared memory (each thread only touches its own row)
#pragma unroll
            for (int c4 = 0; c4 < TS / 4; ++c4) {
                uint4 q;
                *reinterpret_cast<CT*>(&q.x) = v[c4 * 4 + 0] * gs;
                *reinterpret_cast<CT*>(&q.y) = v[c4 * 4 + 1] * gs;
                *reinterpret_cast<CT*>(&q.z) = v[c4 * 4 + 2] * gs;
                *reinterpret_cast<CT*>(&q.w) = v[c4 * 4 + 3] * gs;
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

// EPI: pass 2 with the pointwise epilogue (FusedParams::epilogue) -- its own instantiation, so that the plain pass 2 does
// not carry the branch (measured: 2 us per 8192^2 image when the epilogue was a run-time flag of the one kernel)
template <typename CT, int R, int TS, int MODE, bool RAGGED, bool EPI = false>
__global__ void __launch_bounds__(TS, (TS == 128 ? 3 : 6))
fused_tile_kernel(const __grid_constant__ FusedParams<CT, R> p, const __grid_constant__ CUtensorMap tm_in,
                  const __grid_constant__ CUtensorMap tm_out)
{
    const int64_t b = p.reverse ? (int64_t)(gridDim.x - 1 - blockIdx.x) : (int64_t)blockIdx.x;
    fused_tile_body<CT, R, TS, MODE, RAGGED, false, EPI>(p, tm_in, tm_out, b, FStreamWait{});
}

// ---------------------------------------------------------------------------------------------
// carry chain along one dimension, all scans of that dimension in one launch (fp32 / u32 ring)
//   tail'  = T[s][j] + sum_{q<s} M[q->s] * c_q[j]
//   tau[j] = tail' + P[s] * c_s[j];   c_s[next tile in scan order] = tau[j]
// blockDim = (32 lines, nseg segments of L tiles).  A thread owns L memory-adjacent tiles of one
// line for every scan (so it can re-read the carries of earlier scans it wrote itself); every
// global load of a scan is independent of the recurrence and issued up front, the tables live
// in shared memory, segment tails are exchanged through shared memory.
// ---------------------------------------------------------------------------------------------

template <typename CT, int R>
__device__ __forceinline__ void fmatvec_acc(CT (&y)[R], const CT* m, const CT (&x)[R])
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        CT acc = y[k];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) acc = fmadd(m[k * R + kk], x[kk], acc);
        y[k] = acc;
    }
}

/*
 * Difference basis.  A history is R consecutive outputs of a (usually low-pass) filter: nearly
 * equal numbers, multiplied in the carry algebra by matrices whose rows are large and
 * alternating -- in fp32 the products cancel catastrophically.  The algebra therefore runs on
 * backward differences  (c0, c0-c1, (c0-c1)-(c1-c2), ...)  = D c.  Differences of neighbouring
 * fp32 values are exact (or nearly), the conjugated matrices D M D^-1 (built on the host in
 * fp64, plan.cu) have no cancelling rows, and for integer rings D is unimodular, so the
 * summed-area tables stay bit exact.
 */
template <typename CT, int R>
__device__ __forceinline__ void fdiff_fwd(CT (&c)[R])
{
#pragma unroll
    for (int m = 1; m < R; ++m)
#pragma unroll
        for (int k = R - 1; k >= m; --k) c[k] = c[k - 1] - c[k];
}
template <typename CT, int R>
__device__ __forceinline__ void fdiff_inv(CT (&c)[R])
{
#pragma unroll
    for (int m = R - 1; m >= 1; --m)
#pragma unroll
        for (int k = m; k < R; ++k) c[k] = c[k - 1] - c[k];
}

/*
 * fchain_kernel -- carries of every scan of one dimension, one launch.
 *
 * The kernel is a chain of dependent steps executed once per warp by few warps, so it is bound
 * by instruction issue: it is written for a SMALL instruction footprint and few instructions
 * per tile (a fully unrolled, register-resident version spent most of its time in
 * instruction-cache misses).  The loop over tiles stays rolled and the per-tile data of a thread
 * lives in shared memory, R consecutive words per (tile, thread) so that every access is a
 * pointer plus an immediate:
 *   sT  every scan's tails, brought in by cp.async (no registers, no scoreboard stall) one scan ahead
 *   sC  (= sT, in place) the completed carries of every scan (difference basis), re-used by the
 *       same-dimension residual of the later scans and by the second sweep
 *   sA  (x chain) the A matrices of the block's tile row, see fcrossA_kernel.
 * The matrices of the current tile variant are held in registers and reloaded only when the
 * variant changes (first / last tile of the line).  Offsets are 32-bit (checked by the planner).
 */
template <typename CT, int R>
__device__ __forceinline__ void fload_mat(CT (&m)[R * R], const CT* src)
{
#pragma unroll
    for (int i = 0; i < R * R; ++i) m[i] = src[i];
}
template <typename CT, int R>
__device__ __forceinline__ void fmatvec_acc_reg(CT (&y)[R], const CT (&m)[R * R], const CT (&x)[R])
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        CT acc = y[k];
#pragma unroll
        for (int kk = 0; kk < R; ++kk) acc = fmadd(m[k * R + kk], x[kk], acc);
        y[k] = acc;
    }
}

// MAXT: block size the instantiation is compiled for (256: up to 8 segments, three blocks per SM by registers;
// 512: up to 16 segments)
template <typename CT, int R, int S, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 256 ? (R <= 4 ? 3 : 2) : 1)
fchain_kernel(const __grid_constant__ FChainParams<CT, R> p)
{
    extern __shared__ __align__(16) unsigned char fchain_smem[];
    constexpr int RR = R * R;
    constexpr int NQ = S > 1 ? S - 1 : 1;
    const int L = p.L, nseg = p.nseg;
    const int lane = threadIdx.x, g = threadIdx.y;
    const int nthr = 32 * nseg, tid = g * 32 + lane;
    CT* sA      = reinterpret_cast<CT*>(fchain_smem);          // [nb][S][R][sdk]  (16-byte aligned rows)
    CT* sP      = sA + (p.A ? (size_t)p.nb * S * R * p.sdk : 0); // [V][S][R][R]
    CT* sM      = sP + V_COUNT * S * RR;                        // [V][S][S][R][R]
    CT* sPseg   = sM + V_COUNT * S * S * RR;                    // [S][nseg][R][R]
    CT* segtail = sPseg + S * nseg * RR;                        // [nseg][R][32]
    // per-tile work arrays: slab m holds the R-vector of every thread's m-th tile.  Thread (g, lane) sits at word
    // (g*32 + lane) * RS + g*8 of a slab, RS odd, and slabs are `slot` = 1 (mod 32) words apart: a thread's own
    // accesses, and the panel copies of the strided layout (consecutive tiles = consecutive m, then g), are both
    // free of bank conflicts
    constexpr int RS = fchain_rs(R);
    const int slot = fchain_slot_words(nseg, R);
    CT* sT      = segtail + nseg * R * 32;                      // [S][L][slot]
    CT* sC      = sT;                                           // [S][L][slot]: the carries of a scan overwrite its tails in place

    const int64_t l = p.l0 + (int64_t)blockIdx.x * 32 + lane;
    const bool valid = l < p.l1;
    const int64_t lc = valid ? l : p.l1 - 1;                 // clamp: keep the barriers uniform
    const int j0 = g * L;
    const int cnt = min(p.nb, j0 + L) - j0;                  // tiles of this thread (>= 1)
    const uint32_t plane32 = (uint32_t)p.nb * (uint32_t)p.nl;
    const uint32_t nl32 = (uint32_t)p.sJ;                    // offset step from one tile of a line to the next
    CT* const myT = sT + tid * RS + g * 8;
    CT* const myC = sC + tid * RS + g * 8;
    const CT* const Tl = p.T + lc * p.sL;
    CT* const Cl = p.C + lc * p.sL;

    // strided layout of the signal hierarchy (tile index contiguous, S == 1): the tiles of the block's 32 lines
    // are one contiguous panel per k, moved between global and shared memory by consecutive threads
    // (coalesced) instead of by the thread that owns them
    const bool panel = S == 1 && p.sJ == 1 && p.sL == p.nb;
    const int npanel = 32 * p.nb;
    const int64_t pan0 = (int64_t)blockIdx.x * npanel, panmax = p.nl * p.nb;
    auto panel_slot = [&](int e) -> int {                    // word offset of element e's R-vector in sT / sC
        const int line = e / p.nb, j = e - line * p.nb;
        const int gg = j / L, m = j - gg * L;
        return m * slot + (gg * 32 + line) * RS + gg * 8;
    };
    auto fetch_tails = [&](int s) {                          // tails of scan s -> sT, asynchronously
        if (panel) {
            for (int e = tid; e < npanel; e += nthr)
                if (pan0 + e < panmax) {
                    CT* dst = sT + panel_slot(e);
#pragma unroll
                    for (int k = 0; k < R; ++k) cp_async4(dst + k, p.T + ((int64_t)k * plane32 + pan0 + e));
                }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        uint32_t off = (uint32_t)s * R * plane32 + (uint32_t)j0 * nl32;
        CT* dst = myT + s * L * slot;
        for (int m = 0; m < cnt; ++m, off += nl32, dst += slot)
#pragma unroll
            for (int k = 0; k < R; ++k) cp_async4(dst + k, Tl + (off + k * plane32));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // static tables first (they do not depend on the previous kernel) ...
    pdl_launch_dependents();
    for (int i = tid; i < V_COUNT * S * RR; i += nthr) sP[i] = p.P[i];
    for (int i = tid; i < V_COUNT * S * S * RR; i += nthr) sM[i] = p.M[i];
    for (int i = tid; i < S * nseg * RR; i += nthr) sPseg[i] = p.Pseg[i];
    // x chain with the cross-dimension residual folded in: this line is row `row` of tile row bd
    CT grow[FMAX_SCANS * R];
#pragma unroll
    for (int n = 0; n < FMAX_SCANS * R; ++n) grow[n] = (CT)0;
    const int sdk4 = p.A ? (p.sdk >> 2) : 0;
    if (p.A) {
        const int64_t o = lc / p.Nd;
        const int rem = (int)(lc - o * p.Nd);
        const int bd = rem / p.ts, row = rem - bd * p.ts;
        const int vd = ftile_variant(bd, p.nbd);
#pragma unroll
        for (int n = 0; n < FMAX_SCANS * R; ++n)
            if (n < p.Sd * R) grow[n] = __ldg(p.G + (((int64_t)vd * p.Sd + n / R) * p.ts + row) * R + n % R);
    }
    // ... then everything the previous kernel produced
    pdl_wait();
    if (p.A) {
        // every line of the block lies in the same tile row: stage its A matrices (16 bytes per copy)
        const int64_t o0 = (p.l0 + (int64_t)blockIdx.x * 32) / p.Nd;
        const int bd0 = (int)(((p.l0 + (int64_t)blockIdx.x * 32) - o0 * p.Nd) / p.ts);
        const CT* Arow = p.A + (o0 * p.nbd + bd0) * (int64_t)p.nb * S * R * p.sdk;
        const int n16 = p.nb * S * R * sdk4;
        for (int i = tid; i < n16; i += nthr)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(sA + i * 4)), "l"(Arow + (int64_t)i * 4) : "memory");
    }
    fetch_tails(0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

#pragma unroll
    for (int s = 0; s < S; ++s) {
        const bool causal = p.causal[s] != 0;
        const int gs = causal ? g : nseg - 1 - g;            // segment index in scan order
        const int m0 = causal ? 0 : cnt - 1, dm = causal ? 1 : -1;
        const int dslot = dm * slot;
        CT tau[R];
#pragma unroll
        for (int k = 0; k < R; ++k)
            tau[k] = (gs == 0 && p.ext) ? p.ext[((int64_t)s * R + k) * p.nl + lc] : (CT)0;
        fdiff_fwd<CT, R>(tau);
        if (s > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");

        CT Pm[RR];
        CT Mm[NQ][RR];
        int cur = -1;
        // every tile of this thread is an interior tile (warp-uniform): no variant bookkeeping in the sweeps
        const bool allint = p.uniform || (j0 > 0 && j0 + cnt < p.nb);
        // ---- first sweep, scan order: provisional carries (zero carry into the segment) ----
        {
            int j = j0 + m0;
            const CT* tp = myT + (s * L + m0) * slot;      // == cp: a tile's tail is read before its carry is written
            CT* cp = myC + (s * L + m0) * slot;
            const CT* ap = sA + ((size_t)j * S + s) * R * p.sdk;
            const int da = dm * S * R * p.sdk;
            auto load_mats = [&](int var) {
                fload_mat<CT, R>(Pm, sP + (var * S + s) * RR);
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (q < s) fload_mat<CT, R>(Mm[q], sM + ((var * S + q) * S + s) * RR);
            };
            if (allint) load_mats(V_INTERIOR);
            auto sweep1 = [&](auto interior) {
            for (int t = 0; t < cnt; ++t, j += dm, tp += dslot, cp += dslot, ap += da) {
                if constexpr (!decltype(interior)::value) {
                    const int var = ftile_variant(j, p.nb);
                    if (var != cur) { cur = var; load_mats(var); }   // warp-uniform, at most three times per sweep
                }
                CT x[R];
#pragma unroll
                for (int k = 0; k < R; ++k) x[k] = tp[k];
                if (p.A) {                                    // tail += G_row * A[tile]
#pragma unroll
                    for (int kx = 0; kx < R; ++kx) {
                        CT acc = x[kx];
#pragma unroll
                        for (int i4 = 0; i4 < (FMAX_SCANS * R + 3) / 4; ++i4) {
                            if (i4 < sdk4) {
                                const uint4 qv = *reinterpret_cast<const uint4*>(ap + kx * p.sdk + i4 * 4);
                                const CT a4[4] = { *reinterpret_cast<const CT*>(&qv.x), *reinterpret_cast<const CT*>(&qv.y),
                                                   *reinterpret_cast<const CT*>(&qv.z), *reinterpret_cast<const CT*>(&qv.w) };
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (i4 * 4 + e < FMAX_SCANS * R) acc = fmadd(grow[(i4 * 4 + e) % (FMAX_SCANS * R)], a4[e], acc);
                            }
                        }
                        x[kx] = acc;
                    }
                }
                fdiff_fwd<CT, R>(x);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {                // same-dimension residual of the earlier scans
                    if (q < s) {
                        CT c[R];
                        const CT* cqp = cp - (s - q) * L * slot;
#pragma unroll
                        for (int k = 0; k < R; ++k) c[k] = cqp[k];
                        fmatvec_acc_reg<CT, R>(x, Mm[q], c);
                    }
                }
#pragma unroll
                for (int k = 0; k < R; ++k) cp[k] = tau[k];
                fmatvec_acc_reg<CT, R>(x, Pm, tau);
#pragma unroll
                for (int k = 0; k < R; ++k) tau[k] = x[k];
            }
            };
            if (allint) sweep1(std::true_type()); else sweep1(std::false_type());
        }
        if (s + 1 < S) fetch_tails(s + 1);                   // sT entries of this thread are consumed

        // ---- exchange segment tails, prefix over the earlier segments ----
#pragma unroll
        for (int k = 0; k < R; ++k) segtail[(gs * R + k) * 32 + lane] = tau[k];
        __syncthreads();
        CT u[R];
#pragma unroll
        for (int k = 0; k < R; ++k) u[k] = (CT)0;
        for (int g2 = 0; g2 < gs; ++g2) {
            CT nu[R];
#pragma unroll
            for (int k = 0; k < R; ++k) nu[k] = segtail[(g2 * R + k) * 32 + lane];
            fmatvec_acc<CT, R>(nu, sPseg + (s * nseg + g2) * RR, u);
#pragma unroll
            for (int k = 0; k < R; ++k) u[k] = nu[k];
        }
        if (p.tail_out && gs == nseg - 1 && valid) {
            CT fin[R];
#pragma unroll
            for (int k = 0; k < R; ++k) fin[k] = tau[k];
            fmatvec_acc<CT, R>(fin, sPseg + (s * nseg + gs) * RR, u);
            fdiff_inv<CT, R>(fin);
#pragma unroll
            for (int k = 0; k < R; ++k) p.tail_out[((int64_t)s * R + k) * p.nl + l] = fin[k];
        }
        // ---- second sweep: add the propagated segment carry, store the carries ----
        if (!(S == 1 && p.no_store)) {
            int j = j0 + m0;
            CT* cp = myC + (s * L + m0) * slot;
            uint32_t off = (uint32_t)s * R * plane32 + (uint32_t)j * nl32;
            const uint32_t doff = (uint32_t)dm * nl32;
            cur = -1;
            if (allint) fload_mat<CT, R>(Pm, sP + (V_INTERIOR * S + s) * RR);
            const bool store = valid && !p.no_store;
            auto sweep2 = [&](auto interior) {
            for (int t = 0; t < cnt; ++t, j += dm, cp += dslot, off += doff) {
                if constexpr (!decltype(interior)::value) {
                    const int var = ftile_variant(j, p.nb);
                    if (var != cur) { cur = var; fload_mat<CT, R>(Pm, sP + (var * S + s) * RR); }
                }
                CT c[R];
#pragma unroll
                for (int k = 0; k < R; ++k) c[k] = cp[k] + u[k];
                if (s + 1 < S) {                              // later scans need the completed carries (difference basis)
#pragma unroll
                    for (int k = 0; k < R; ++k) cp[k] = c[k];
                }
                fdiff_inv<CT, R>(c);
                if (panel) {                                  // leaves through shared memory, see below
#pragma unroll
                    for (int k = 0; k < R; ++k) cp[k] = c[k];
                } else if (store) {
#pragma unroll
                    for (int k = 0; k < R; ++k) Cl[off + k * plane32] = c[k];
                }
                CT nu[R];
#pragma unroll
                for (int k = 0; k < R; ++k) nu[k] = (CT)0;
                fmatvec_acc_reg<CT, R>(nu, Pm, u);
#pragma unroll
                for (int k = 0; k < R; ++k) u[k] = nu[k];
            }
            };
            if (allint) sweep2(std::true_type()); else sweep2(std::false_type());
        }
        __syncthreads();                                      // segtail is reused by the next scan
        if (panel && !p.no_store) {
            for (int e = tid; e < npanel; e += nthr)
                if (pan0 + e < panmax) {
                    const CT* src = sC + panel_slot(e);
#pragma unroll
                    for (int k = 0; k < R; ++k) p.C[(int64_t)k * plane32 + pan0 + e] = src[k];
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// carries of a short-memory dimension, no chain.  The serial inter-tile loop (lib/split.cpp:832-846,
// ctail[t] = tail[t] + P * ctail[t-1]) multiplies the carry by the tile's transition matrix P = A^tile.  For a
// filter whose impulse response is much shorter than a tile (the headline Gaussian: |pole| = 0.79, P ~ 6e-13 over
// 128 samples) that product is far below the last bit of the tail it is added to, so the carry entering a tile is
// the tail' of the tile before it -- every (line, tile) pair is independent and the carry stage becomes one
// streaming launch per dimension.  The planner takes this path only when every entry of P (all tile variants,
// difference basis, fp64) is below 1e-10 in magnitude; the result differs from the chained one by less than that.
//   C_0[j] = T'_0[j -/+ 1]
//   C_1[j] = T'_1[j -/+ 1] + M[0 -> 1] * C_0[j -/+ 1]            (same-dimension residual, lib/split.cpp:912-1004)
// with T' = T + G_row * A[tile] in the x dimension (cross-dimension residual, as in fchain_kernel).
// ---------------------------------------------------------------------------------------------
constexpr int FLOCAL_TPT = 8;     // tiles per thread (their loads are issued together: the kernel is pure latency otherwise)

template <typename CT, int R>
__global__ void __launch_bounds__(128)
flocal_kernel(const __grid_constant__ FLocalParams<CT, R> p)
{
    constexpr int TPT = FLOCAL_TPT, HALO = 2, NT = TPT + 2 * HALO;
    const int64_t l = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int j0 = blockIdx.y * TPT;
    pdl_launch_dependents();
    // x dimension: this line is row `row` of tile row bd (static tables: before the wait)
    CT grow[FMAX_SCANS * R];
    int64_t arow = 0;
    const bool lv = l < p.nl;
    if (p.A && lv) {
        const int64_t o = l / p.Nd;
        const int rem = (int)(l - o * p.Nd);
        const int bd = rem / p.ts, row = rem - bd * p.ts;
        const int vd = ftile_variant(bd, p.nbd);
#pragma unroll
        for (int n = 0; n < FMAX_SCANS * R; ++n)
            grow[n] = n < p.Sd * R ? __ldg(p.G + (((int64_t)vd * p.Sd + n / R) * p.ts + row) * R + n % R) : (CT)0;
        arow = (o * p.nbd + bd) * (int64_t)p.nb * p.S * R * p.sdk;
    }
    pdl_wait();
    if (!lv) return;
    const int64_t plane = (int64_t)p.nb * p.nl;
    const int S = p.S;
    // tail' of both scans at the tiles j0 - HALO .. j0 + TPT + HALO - 1 (zero outside the line): every load first
    CT t[2][NT][R];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int tt = j0 - HALO + i;
            const bool in = s < S && tt >= 0 && tt < p.nb;
#pragma unroll
            for (int k = 0; k < R; ++k)
                t[s][i][k] = in ? ld_stream<CT>(p.T + ((int64_t)s * R + k) * plane + (int64_t)tt * p.nl + l) : (CT)0;
        }
    if (p.A) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int tt = j0 - HALO + i;
                if (s < S && tt >= 0 && tt < p.nb) {
                    const CT* ap = p.A + arow + ((int64_t)tt * S + s) * R * p.sdk;
#pragma unroll
                    for (int kx = 0; kx < R; ++kx) {
                        CT acc = t[s][i][kx];
#pragma unroll
                        for (int n = 0; n < FMAX_SCANS * R; ++n)
                            if (n < p.Sd * R) acc = fmadd(grow[n], __ldg(ap + kx * p.sdk + n), acc);
                        t[s][i][kx] = acc;
                    }
                }
            }
    }
    // y = D^-1 (M' (D c)), M' of the tile variant `var`
    auto residual = [&](int var, const CT (&c)[R], CT (&y)[R]) {
        CT d[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { d[k] = c[k]; y[k] = (CT)0; }
        fdiff_fwd<CT, R>(d);
        const CT* m = p.M + (((int64_t)var * S + 0) * S + 1) * R * R;
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int kk = 0; kk < R; ++kk) y[k] = fmadd(__ldg(m + k * R + kk), d[kk], y[k]);
        fdiff_inv<CT, R>(y);
    };
    // (register arrays need compile-time indices: the scan directions select between static neighbours)
    const bool f0 = p.causal[0] != 0, f1 = p.causal[1] != 0;
    const int d1 = f1 ? 1 : -1;
#pragma unroll
    for (int jj = 0; jj < TPT; ++jj) {
        const int j = j0 + jj;
        if (j >= p.nb) break;
        const int i = jj + HALO;                                  // index of tile j in t[][]
        // scan 0: the carry entering tile j is the tail' of the tile before it
        CT c0[R], c0p[R];                                         // ... entering tile j, entering the tile before j in scan 1's order
#pragma unroll
        for (int k = 0; k < R; ++k) {
            c0[k] = f0 ? t[0][i - 1][k] : t[0][i + 1][k];
            c0p[k] = f1 ? (f0 ? t[0][i - 2][k] : t[0][i][k]) : (f0 ? t[0][i][k] : t[0][i + 2][k]);
            p.C[(int64_t)k * plane + (int64_t)j * p.nl + l] = c0[k];
        }
        if (p.tail_out && (p.causal[0] ? j == p.nb - 1 : j == 0)) {
#pragma unroll
            for (int k = 0; k < R; ++k) p.tail_out[(int64_t)k * p.nl + l] = t[0][i][k];
        }
        if (S < 2) continue;
        // scan 1: tail' of the tile before j (in its order) + the same-dimension residual of scan 0's carry there
        const int jp = j - d1;
        CT c1[R];
#pragma unroll
        for (int k = 0; k < R; ++k) c1[k] = f1 ? t[1][i - 1][k] : t[1][i + 1][k];
        if (jp >= 0 && jp < p.nb) {
            CT y[R];
            residual(ftile_variant(jp, p.nb), c0p, y);
#pragma unroll
            for (int k = 0; k < R; ++k) c1[k] = c1[k] + y[k];
        }
#pragma unroll
        for (int k = 0; k < R; ++k) p.C[((int64_t)R + k) * plane + (int64_t)j * p.nl + l] = c1[k];
        if (p.tail_out && (p.causal[1] ? j == p.nb - 1 : j == 0)) {
            CT y[R];
            residual(ftile_variant(j, p.nb), c0, y);
#pragma unroll
            for (int k = 0; k < R; ++k) p.tail_out[((int64_t)R + k) * p.nl + l] = t[1][i][k] + y[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cross-dimension residual (/root/reference/lib/split.cpp:1215-1633), part 1: one warp per tile.
// The completed d carries change the d-filtered tile by G_d * CY (rows x cols, rank R per d
// scan); the x tails P1 took from the incomplete tile therefore miss (G_d * CY) * L_x^T:
//   A[tile][sx][kx][sd*R+k] = sum over the tile's columns of CY_sd[k][col] * L_sx[kx][col]
// Part 2 (TX[sx][kx][row] += sum_sd sum_k G[sd][row][k] * A[..]) is applied by the x chain while
// it loads its tails, so the x tails are never rewritten in memory.
// ---------------------------------------------------------------------------------------------
template <typename CT, int R, int TS, bool TAILS>
__device__ __forceinline__ void fcrossA_body(const FCrossParams<CT, R>& p, const int64_t w)
{
    constexpr int CPL = TS / 32;                 // columns per lane
    const int lane = threadIdx.x & 31;
    int64_t b = w;
    const int bx = (int)(b % p.nbx); b /= p.nbx;
    const int bd = (int)(b % p.nbd);
    const int64_t o = b / p.nbd;
    const int vx = ftile_variant(bx, p.nbx);
    const int64_t ly0 = o * p.Nx + (int64_t)bx * TS + lane;
    const int64_t kstride_y = (int64_t)p.nbd * p.nly;

    // the d carries of this tile's columns, every d scan, in the difference basis (G is stored as G * D^-1)
    CT cy[FMAX_SCANS][R][CPL];
    if (p.local) {
        // short-memory d dimension: the carries are the tails of the tiles above / below (see FusedParams::local)
#pragma unroll
        for (int sd = 0; sd < FMAX_SCANS; ++sd)
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int c = 0; c < CPL; ++c) cy[sd][k][c] = (CT)0;
        const int dd0 = p.causal_d[0] ? 1 : -1, dd1 = p.causal_d[1] ? 1 : -1;
        auto tail = [&](int s, int t, int c, CT (&x)[R]) {
            const bool ok = t >= 0 && t < p.nbd && (int64_t)bx * TS + lane + c * 32 < p.Nx;
#pragma unroll
            for (int k = 0; k < R; ++k) x[k] = ok ? p.TY[((int64_t)s * R * p.nbd + t) * p.nly + k * kstride_y + ly0 + c * 32] : (CT)0;
        };
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            CT x[R];
            tail(0, bd - dd0, c, x);
#pragma unroll
            for (int k = 0; k < R; ++k) cy[0][k][c] = x[k];
            if (p.Sd > 1) {
                const int t1 = bd - dd1;
                tail(1, t1, c, x);
                if (t1 >= 0 && t1 < p.nbd) {
                    CT c0p[R], y[R];
                    tail(0, t1 - dd0, c, c0p);
                    flocal_residual<CT, R>(p.Md + (((int64_t)ftile_variant(t1, p.nbd) * p.Sd + 0) * p.Sd + 1) * R * R, c0p, y);
#pragma unroll
                    for (int k = 0; k < R; ++k) x[k] = x[k] + y[k];
                }
#pragma unroll
                for (int k = 0; k < R; ++k) cy[1][k][c] = x[k];
            }
        }
    } else {
#pragma unroll
    for (int sd = 0; sd < FMAX_SCANS; ++sd)
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                cy[sd][k][c] = (sd < p.Sd && (int64_t)bx * TS + lane + c * 32 < p.Nx)      // columns of a partial tile
                                   ? p.CY[((int64_t)sd * R * p.nbd + bd) * p.nly + k * kstride_y + ly0 + c * 32] : (CT)0;
    }
#pragma unroll
    for (int sd = 0; sd < FMAX_SCANS; ++sd)
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            CT h[R];
#pragma unroll
            for (int k = 0; k < R; ++k) h[k] = cy[sd][k][c];
            fdiff_fwd<CT, R>(h);
#pragma unroll
            for (int k = 0; k < R; ++k) cy[sd][k][c] = h[k];
        }

    CT* Aout = p.A + w * ((int64_t)p.Sx * R * p.sdk);
    // tails mode (short-memory pass 2, FusedParams::local == 2): the x tails of this tile get their cross-dimension
    // residual G_row * A added in place -- pass 2 then reads corrected tails and needs neither A nor G.  This warp owns the
    // tile: its 128 rows are 4 per lane; gv = the G rows of those rows (at most two d scans in this mode)
    constexpr int NL = TAILS ? 2 * R : 1;
    CT gv[CPL][NL];
    constexpr bool fix_tails = TAILS;               // (p.local == 2)
    if constexpr (TAILS) {
        const int vd = ftile_variant(bd, p.nbd);
#pragma unroll
        for (int c = 0; c < CPL; ++c)
#pragma unroll
            for (int n = 0; n < NL; ++n)
                gv[c][n] = n < p.Sd * R ? __ldg(p.G + (((int64_t)vd * p.Sd + n / R) * TS + lane + c * 32) * R + n % R) : (CT)0;
    }
    for (int q = 0; q < p.Sx; ++q) {
        CT lv[R][CPL];
#pragma unroll
        for (int kx = 0; kx < R; ++kx)
#pragma unroll
            for (int c = 0; c < CPL; ++c)
                lv[kx][c] = __ldg(p.L + (((int64_t)vx * p.Sx + q) * R + kx) * TS + c * 32 + lane);
#pragma unroll
        for (int kx = 0; kx < R; ++kx) {
            CT mine = (CT)0;                                 // lane n keeps entry n = sd * R + k
            CT av[NL];                                       // every lane: the whole row A[q][kx][.] (butterfly sums)
#pragma unroll
            for (int n = 0; n < NL; ++n) av[n] = (CT)0;
#pragma unroll
            for (int sd = 0; sd < FMAX_SCANS; ++sd) {
                if (sd < p.Sd) {
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        CT x = (CT)0;
#pragma unroll
                        for (int c = 0; c < CPL; ++c) x = fmadd(cy[sd][k][c], lv[kx][c], x);
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, off);
                        if (lane == sd * R + k) mine = x;
                        if (TAILS && sd < 2) av[(sd * R + k) % NL] = x;
                    }
                }
            }
            if constexpr (!fix_tails) {
                if (lane < p.sdk) Aout[((int64_t)q * R + kx) * p.sdk + lane] = mine;
            } else {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int row = lane + c * 32;
                    if ((int64_t)bd * TS + row < p.Nd) {
                        CT* tx = p.TXw + (((int64_t)q * R + kx) * p.nbx + bx) * p.nlx + o * p.Nd + (int64_t)bd * TS + row;
                        CT acc = *tx;
#pragma unroll
                        for (int n = 0; n < NL; ++n) acc = fmadd(gv[c][n], av[n], acc);
                        *tx = acc;
                    }
                }
            }
        }
    }
}

template <typename CT, int R, int TS, bool TAILS>
__global__ void __launch_bounds__(128)
fcrossA_kernel(const __grid_constant__ FCrossParams<CT, R> p)
{
    const int64_t w = p.w0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (w >= p.w1) return;
    fcrossA_body<CT, R, TS, TAILS>(p, w);
}

// ---------------------------------------------------------------------------------------------
// Both sweeps in ONE launch, for filters with short memory in every scanned dimension (FusedParams::local: the
// carries entering a tile follow from the tails of its neighbours, there is no chain).  The two-sweep scheme reads the
// image twice from HBM because pass 2 of ANY tile has to wait for pass 1 of EVERY tile; here pass 2 of tile row r only
// waits for pass 1 of rows r-2 .. r+2, so it can follow pass 1 at a distance of a few tile rows and find its input
// still in L2: 8 bytes of HBM traffic per sample instead of 12.
//   Work items are handed out by a ticket counter in dependency order -- step s: pass 1 of tile row s, the cross
// residuals A of row s - lag_a (one item = 4 tiles, fcrossA_body), pass 2 of row s - lag_p -- so everything an item
// waits for has an earlier ticket and is resident or finished: no deadlock, whatever order the hardware starts CTAs in.
// Finished items count up per-row counters (release); a waiting item polls them (acquire, one thread, bounded).
// The counters and the ticket are zeroed before each launch (a memset node in front of the kernel).
//   OFF by default (RFB_STREAM=1).  Measured (profiles/r02_stream_*): the unrolled scans are 41 KB (pass 1) and 64 KB
// (pass 2) of code.  In a one-sweep launch the three resident CTAs of an SM start together, take the same time and so run
// through that code in step, sharing the instruction fetch (hit rate 90-95 %); here an SM holds a mix of items at
// unrelated phases and the instruction cache thrashes (hit rate 71 %, `no_instruction` 2.0 stalls per issue against
// 0.3-0.5) -- also with ONE body for both passes told apart at run time (74 %), and with the SMs split into a pass-1
// and a pass-2 class (81 %).  8192^2 therefore takes 250 us per image instead of 156 although pass 2 can find its
// input in L2; stacks that fit L2 anyway tie with the two sweeps (4096^2: 79.2 against 78.9 us).
// ---------------------------------------------------------------------------------------------
template <typename CT, int R, int TS>
__global__ void __launch_bounds__(TS, 3)
fused_stream_kernel(const __grid_constant__ FStreamParams<CT, R> sp, const __grid_constant__ CUtensorMap tm_in,
                    const __grid_constant__ CUtensorMap tm_out)
{
    // the ticket is broadcast through the spare word beside the tile's mbarrier (no static shared memory: three CTAs
    // per SM leave ~1 KB)
    extern __shared__ __align__(16) unsigned char fsmem_raw[];
    unsigned char* tile0 = fsmem_raw + ((1024u - (smem_u32(fsmem_raw) & 1023u)) & 1023u);
    volatile unsigned* s_ticket = reinterpret_cast<volatile unsigned*>(tile0 + (TS / 32) * TS * 128 + 8);
    if (threadIdx.x == 0) *s_ticket = atomicAdd(sp.ticket, 1u);
    __syncthreads();
    const unsigned t = *s_ticket;
    const int s = (int)(t / (unsigned)sp.step), q = (int)(t % (unsigned)sp.step);
    const int nbx = sp.t.nbx;
    auto done = [&](unsigned* counter) {
        __syncthreads();                                   // every thread's tails / residuals are written ...
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(counter, 1u); }      // ... and visible before the count
    };
    FStreamWait sy;
    sy.cnt_p1 = sp.cnt_p1; sy.cnt_a = sp.cnt_a; sy.err = sp.err; sy.rows = sp.rows;
    sy.need_p1 = (unsigned)nbx; sy.need_a = 0; sy.row = 0;
    if (q >= nbx && q < nbx + sp.na) {
        const int row = s - sp.lag_a;
        if (row < 0 || row >= sp.rows) return;
        sy.row = row;
        if (threadIdx.x == 0) fstream_wait(sy);
        __syncthreads();
        const int tx = (q - nbx) * (TS / 32) + (threadIdx.x >> 5);              // one warp per tile
        if (tx < nbx) fcrossA_body<CT, R, TS, false>(sp.c, (int64_t)row * nbx + tx);
        done(sp.cnt_a + row);
        return;
    }
    if (q < nbx) {
        const int row = s;
        if (row >= sp.rows) return;
        fused_tile_body<CT, R, TS, FMODE_P1, false, true>(sp.t, tm_in, tm_in, (int64_t)row * nbx + q, sy);
        done(sp.cnt_p1 + row);
    } else {
        const int row = s - sp.lag_p;
        if (row < 0 || row >= sp.rows) return;
        sy.row = row; sy.need_a = (unsigned)sp.na;
        fused_tile_body<CT, R, TS, FMODE_P2, false, true>(sp.t, tm_in, tm_out, (int64_t)row * nbx + (q - nbx - sp.na), sy);
    }
}

} // namespace rfb
