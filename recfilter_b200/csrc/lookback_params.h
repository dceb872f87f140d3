#pragma once
/*
 * lookback_params.h -- kernel parameter blocks of the single-pass (decoupled look-back) kernels,
 * shared by the launch planner (plan.cu) and the kernels (lookback.cuh).
 */
#include <stdint.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include "engine.h"

namespace rfb {

// tag of a published chunk: (epoch << 2) | state (see lookback.cuh).  The epoch changes with every launch, so the
// records are never cleared: a chunk of an older epoch reads as "nothing published".
enum { LB_NONE = 0, LB_AGGREGATE = 1, LB_INCLUSIVE = 2 };

constexpr int LB_SIGNAL_REC_CHUNKS = 4;           // 64 bytes per tile (3 chunks used at order 8)
constexpr uint32_t LB_SPIN_LIMIT = 1u << 22;     // polls before a CTA gives up and raises the error flag (no hang)

// one dimension of the 2-D look-back kernel: at most one scan
template <typename CT, int R>
struct LBDim {
    int nscan;                    // 0: no scan along this dimension
    int causal;
    CT  a[R + 1];                 // a[0]: clamp-history factor (1/b0), a[1..R]: feedback (unit feed-forward form)
    CT  P[R * R];                 // response of a tile's tail to the carry entering it (difference basis)
    const CT* Ppow;               // [nb][R][R]: P^j, j = 0 .. nb-1 (difference basis)
    void* rec;                    // [tile][line] records of (R + 2) / 3 16-byte chunks {3 values, tag}: the line's
                                  // aggregate (tail of the tile scanned with zero history), later overwritten by its
                                  // inclusive (completed tail = the carry entering the next tile); difference basis
};

template <typename CT, int R>
struct LBTileParams {
    int64_t Nx, Nd, No;
    int nbx, nbd;
    int clamp;
    CT  gain;                     // product of the feed-forward coefficients, applied at the store
    int prefetch;                 // > 0: L2 prefetch of the tile `prefetch` tickets ahead
    int rows_first;               // 1: tiles are handed out row-major instead of along anti-diagonals
    unsigned long long* ticket;   // ticket counter, never reset: ticket / tiles = launch number (epoch of the tags), % tiles = tile
    uint32_t* err;                // set to 1 if a CTA ran into LB_SPIN_LIMIT
    LBDim<CT, R> x, d;
};

// long 1-D signals: rows of TS samples, TS rows per CTA, one scan
template <typename CT, int R>
struct LBSignalParams {
    int64_t rows;                 // rows in total (signals x rows per signal)
    int tile_rows;                // rows per CTA: 128, 64 or 32
    int tiles_per_signal;         // CTAs per signal
    int causal;
    int clamp;
    CT  gain;
    CT  a[R + 1];
    CT  Pstep[5][R * R];          // P^(1,2,4,8,16): row transitions of the intra-warp scan (difference basis)
    CT  Pwarp[R * R];             // P^32
    CT  Q[R * R];                 // P^tile_rows: transition of a whole tile
    CT  Q32[R * R];               // Q^32: one look-back window
    const CT* Plane;              // [R*R][32]: P^lane
    const CT* Qpow;               // [R*R][32]: Q^k
    void* rec;                    // [tile] records of LB_SIGNAL_REC_CHUNKS 16-byte chunks ((R + 2) / 3 used): aggregate, then inclusive
    int prefetch;                 // > 0: L2 prefetch of the tile `prefetch` tickets ahead
    int pass0_first_chunk;        // short-memory filters: the first chunks (32 samples each) of a row do not reach its tail
    int depth1;                   // short memory across a whole tile: the carry is the previous tile's aggregate
    unsigned long long* ticket;   // see LBTileParams
    uint32_t* err;
};

inline size_t lb_tile_smem_bytes(int ts) { return (size_t)ts * ts * 4 + 1024 + 64; }

} // namespace rfb
