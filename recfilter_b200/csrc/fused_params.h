#pragma once
/*
 * fused_params.h -- kernel parameter blocks of the fused fast path (shared by the launch
 * planner in plan.cu and the kernels in fused.cuh).  See fused.cuh for the algorithm.
 */
#include <stdint.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include "engine.h"

namespace rfb {

constexpr int FMAX_SCANS = 4;     // scans per dimension handled by the fused path
constexpr int FCHAIN_L   = 8;     // tiles per thread in the carry chain

enum { FMODE_P1 = 0, FMODE_P2 = 1 };

template <typename CT, int R>
struct FusedScanTab {
    int causal[FMAX_SCANS];
    CT  a[FMAX_SCANS][R + 1];     // a[s][0]: clamp-history factor (1/b0); a[s][1..R]: feedback
};

template <typename CT, int R>
struct FusedParams {
    int64_t Nx, Nd, No;
    int64_t o0;                   // first outer index of this launch (a slice of a stack) ...
    int64_t No_launch;            // ... and the images its grid covers (0: all No)
    int nbx, nbd;
    int clamp;
    int x_lo_closed, x_hi_closed, d_lo_closed, d_hi_closed;
    int mx, md;
    int reverse;                  // walk the tiles backwards (pass 2: most recently read first)
    int prefetch;                 // > 0: a CTA prefetches (L2) the tile of block blockIdx + prefetch
    CT  gain;                     // product of the feed-forward coefficients
    CT* TX; const CT* CX;         // x tails (out, P1) / carries (in, P2)   [s][k][bx][lx]
    CT* TY; const CT* CY;         // d tails / carries                     [s][k][bd][ly]
    int64_t nlx, nly;
    // signal mode (long 1-D signals): the array is viewed as rows of one tile width; a row continues the
    // row before it, sig_rows rows make one signal (only its first / last row is closed), md = 0, nbx = 1
    int signal;
    int64_t sig_rows;
    // pointwise epilogue fused into the store of pass 2 (epilogue != 0; needs x scans): out = epi_out * filtered + epi_in * input
    // -- the unsharp mask (1+w)*image - w*blur of apps/usm/unsharp_mask_optimized.cpp:61-66, which the reference merges
    // with RecFilter::compute_at(USM, ...).  The input tile is fetched again (TMA, an L2 hit) behind the row scans.
    int epilogue;
    CT  epi_in, epi_out;
    // short-memory pass 2 (local != 0; at most two scans per dimension): the carries entering the tile are derived
    // from the TAILS of the neighbouring tiles on the fly -- no carry kernels ran, CX / CY are not read:
    //   C_0 = T'_0[tile before],  C_1 = T'_1[tile before] + M[0 -> 1] * C_0[there],  T' = T + G_row * A[tile] along x
    int local;
    const CT* Mx; const CT* Md;   // [V][S][S][R][R]  same-dimension residual, difference basis
    CT Mlx[V_COUNT][R * R], Mld[V_COUNT][R * R];   // M[0 -> 1] per tile variant, as kernel constants (no load latency in the prologue)
    const CT* A;                  // [tile][sx][kx][sdk]  cross-dimension residual (fcrossA_kernel) or null
    const CT* G;                  // [V][Sd][ts][R]
    int sdk;
    FusedScanTab<CT, R> sx, sd;
};

// The carry algebra of the fused path runs in the compute type (fp32 / the u32 ring): the tables
// are generated on the host in fp64 and rounded once.
template <typename CT, int R>
struct FChainParams {
    const CT* T; CT* C;           // [s][k][j][l]
    int64_t nl; int nb; int S;
    int64_t l0, l1;               // lines [l0, l1) are chained by this launch (l1 == 0: all nl lines); l0 a multiple of 32
    int nseg;                     // segments of L tiles (threadIdx.y)
    int L;                        // tiles per thread (4 or FCHAIN_L)
    int causal[FMAX_SCANS];
    const CT* P;                  // [V][S][R][R]
    const CT* M;                  // [V][S][S][R][R]   (q -> s)
    const CT* Pseg;               // [S][nseg][R][R]   product over a segment, segments in scan order
    const CT* ext;                // [s][k][l] carry entering the first tile (shard cut) or null
    CT* tail_out;                 // [s][k][l] completed tail leaving the last tile or null
    // addressing of T / C: element (s, k, tile j, line l) at (s*R + k) * nb*nl + j * sJ + l * sL
    int64_t sJ, sL;               // default layout ("line fastest"): sJ = nl, sL = 1
    int uniform;                  // 1: every tile is an interior tile (signal hierarchy)
    int no_store;                 // 1: only tail_out is wanted (up-sweep of the signal hierarchy)
    // x chain only: cross-dimension residual applied while the tails are loaded (null: plain chain)
    const CT* A;                  // [tile][s][kx][sdk]  from fcrossA_kernel
    const CT* G;                  // [V][Sd][ts][R]      d response to a unit carry (times D^-1)
    int64_t Nd; int nbd; int Sd; int ts; int sdk;
};

// carries of a SHORT-MEMORY dimension (see flocal_kernel in fused.cuh): at most two scans
template <typename CT, int R>
struct FLocalParams {
    const CT* T; CT* C;           // [s][k][j][l], as the chain kernel
    int64_t nl; int nb; int S;
    int causal[2];
    const CT* M;                  // [V][S][S][R][R]   (q -> s), difference basis
    CT* tail_out;                 // [s][k][l] completed tail leaving the last tile or null
    // x dimension only: cross-dimension residual applied while the tails are read (null: none), as in FChainParams
    const CT* A; const CT* G;
    int64_t Nd; int nbd; int Sd; int ts; int sdk;
};

template <typename CT, int R>
struct FCrossParams {
    const CT* CY;                 // [sd][k][bd][ly]   completed d carries
    CT* A;                        // [tile][sx][kx][sdk]  sdk = Sd * R rounded up to a multiple of 4
    const CT* L;                  // [V][Sx][R][TS]
    int64_t Nx, Nd, No; int nbx, nbd; int Sx, Sd; int sdk; int64_t nly, nlx;
    int64_t w0, w1;               // tiles [w0, w1) are handled by this launch (w1 == 0: all of them)
    // short-memory d dimension (local != 0): CY is not read; the d carries are derived from the d tails TY of the
    // tiles above / below on the fly (see FusedParams::local)
    int local;                    // 1: A is stored (pass 2 applies G_row * A itself); 2: A is applied to the x tails TXw in place
    const CT* TY; const CT* Md;
    int causal_d[2];
    CT* TXw; const CT* G;
};

// dynamic shared memory of one chain block (layout in fchain_kernel)
// shared-memory geometry of the chain kernel's per-tile work arrays (see fchain_kernel)
__host__ __device__ constexpr int fchain_rs(int R) { return (R % 2 == 0) ? R + 1 : R; }
__host__ __device__ inline int fchain_slot_words(int nseg, int R)
{
    return ((32 * nseg * fchain_rs(R) + nseg * 8 + 31) / 32) * 32 + 1;
}
inline size_t fchain_smem_bytes(int S, int nseg, int R, int L, int nb, int sdk_if_cross)
{
    // the carries of a scan overwrite its tails in place
    const size_t work = (size_t)L * fchain_slot_words(nseg, R) * S;
    return ((size_t)nb * S * R * sdk_if_cross + (size_t)V_COUNT * S * R * R + (size_t)V_COUNT * S * S * R * R +
            (size_t)S * nseg * R * R + (size_t)nseg * R * 32 + work) * 4;
}

// both sweeps in one launch (fused_stream_kernel, fused.cuh)
template <typename CT, int R>
struct FStreamParams {
    FusedParams<CT, R> t;           // the tile items: pass 1 and the short-memory pass 2 (local = 1)
    FCrossParams<CT, R> c;          // the cross-residual items (local = 1)
    unsigned* ticket;               // device words, zeroed before every launch: [ticket][pad][cnt_p1 rows][cnt_a rows]
    unsigned* cnt_p1;               // pass-1 tiles finished per tile row
    unsigned* cnt_a;                // A items finished per tile row
    unsigned* err;                  // set (never cleared by the kernel) when an item gave up waiting
    int rows;                       // tile rows of the whole stack: No * nbd
    int na;                         // A items per tile row: ceil(nbx / 4)
    int step;                       // tickets per step: nbx + na + nbx
    int lag_a, lag_p;               // tile rows the cross residuals / pass 2 run behind pass 1
};

// dynamic shared memory of one tile CTA: the swizzled boxes, alignment slack, the mbarrier
// (pass 2 adds the staged carries of the tile: nscans * R * ts words)
// staged carries of a pass-2 CTA: one slot of R x ts words per scan; the short-memory variant adds one temporary per
// dimension and the A matrices of three x neighbours
inline int fused_p2_carry_words(int mx, int md, int R, int ts, int local, int sdk)
{
    return local ? (mx + md + 2) * R * ts + 3 * mx * R * sdk : (mx + md) * R * ts;
}
inline size_t fused_tile_smem_bytes(int ts, int carry_words = 0) { return (size_t)ts * ts * 4 + 1024 + 16 + (size_t)carry_words * 4; }

// TMA descriptor of a dense [rows][Nx] matrix of 4-byte elements, box = 32 columns x ts rows, 128 B swizzle
// (defined once in plan.cu; resolves cuTensorMapEncodeTiled through the runtime, no libcuda link)
cudaError_t make_tile_map(CUtensorMap* map, const void* base, int64_t Nx, int64_t rows, int ts, bool is_float);

} // namespace rfb
