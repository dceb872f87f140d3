/*
 * engine.h -- internal data model shared by the launch planner (plan.cpp), the
 * kernels (kernels.cu) and the C ABI (capi.cpp).  Not part of the public boundary.
 *
 * Vocabulary (follows the reference, lib/recfilter_internals.h:23-44 and
 * lib/split.cpp:35-61):
 *   scan     one add_filter() call: order-r causal/anticausal recurrence along a dim
 *   tile     T consecutive samples of a line that one thread scans in registers
 *   tail     last r outputs of a scan inside a tile, in scan order (k=0 newest)
 *   carry    the completed tail of the previous tile = history entering a tile
 *   pass     one read+write sweep over the array that applies all scans of one
 *            dimension (or of dimensions 0 and 1 fused)
 *
 * A pass sees the array as a dense 3-level view [No][Nd][Nx]: Nx is contiguous
 * ("row" dimension, scans along it need the shared-memory transpose), Nd is the
 * "column" dimension (stride Nx) and No is everything outside.
 */
#pragma once
#include <stdint.h>
#include <vector>
#include <string>
#include "../../include/recfilter_b200.h"

namespace rfb {

constexpr int TILE = 64;          // register tile: samples per thread per scan
constexpr int MAX_SCANS_DIM = 32; // scans along one dimension in one pass
constexpr int RFB_MAX_DEVICES = 64; // per-device caches of function attributes

// tile position classes along one dimension
enum Variant { V_FIRST = 0, V_INTERIOR = 1, V_LAST = 2, V_SINGLE = 3, V_COUNT = 4 };

// geometry of one dimension of a pass
struct DimGeom {
    int64_t n = 1;        // extent
    int     t = TILE;     // logical tile length (<= TILE)
    int     nb = 1;       // number of tiles
    int     len_last = 0; // length of the last tile
    int     lo_closed = 1, hi_closed = 1;   // 0 when the face is a shard cut
    int     nscans = 0;
};

// accumulation / table type of the carry algebra: the small K2/K3 kernels work in fp64 for
// float filters (independent rounding errors of the r history values are amplified by the
// recurrence, so carries must be much better than fp32-accurate before they are rounded once)
template <typename CT> struct TabType { typedef CT type; };
#ifndef RFB_FP32_CARRY_ALGEBRA
template <> struct TabType<float> { typedef double type; };
#endif

// device-side scan table entry (coefficients already converted to the compute type)
template <typename CT, int R>
struct ScanTab {
    int causal[MAX_SCANS_DIM];
    CT  coef[MAX_SCANS_DIM][R + 1];
};

// kernel parameters of one pass (passed by value)
template <typename CT, int R>
struct PassParams {
    // view
    int64_t Nx, Nd, No;
    int tx, td;            // logical tile lengths
    int nbx, nbd;          // tiles per dimension
    int lenx_last, lend_last;
    int signal_mode;       // 1: rows of a CTA are consecutive tiles of one line
    int clamp;             // clamped image border
    int x_lo_closed, x_hi_closed, d_lo_closed, d_hi_closed;
    int mx, md;            // scans along x / along d
    // carry storage (compute type)
    CT* TX; CT* CX;        // x tails / carries   [s][k][bx][lx]  (signal mode: [s][k][lx][bx])
    CT* TY; CT* CY;        // d tails / carries   [s][k][bd][ly]
    int64_t nlx, nly;      // number of x lines (No*Nd) and d lines (No*Nx)
    ScanTab<CT, R> sx, sd;
};

// parameters of the carry-chain kernels for one scan of one dimension
template <typename CT, int R>
struct ChainParams {
    typedef typename TabType<CT>::type TT;
    CT* T; CT* C;               // tails in, carries out
    int64_t nl;                 // lines
    int nb;                     // tiles
    int64_t tile_stride, line_stride, plane;   // addressing: (s*R+k)*plane + j*tile_stride + l*line_stride
    int s;                      // scan being chained
    int causal;
    int seg, nseg;              // tiles per segment, segments
    const TT* P;                // [V_COUNT][S][R][R]
    const TT* M;                // [V_COUNT][S][S][R][R]  (q -> s)
    int S;                      // scans in this dimension
    const TT* Pseg;             // [2][R][R]: product over a full interior segment / over the last segment
    TT* SEGT; TT* SEGC;         // [k][g][l] segment tails / carries
    const CT* ext;              // external carry-in for this scan [k][l] or null
    CT* tail_out;               // final outgoing tail [k][l] or null (sharding)
};

// parameters of the cross-dimension residual kernel (x carries -> d tails)
template <typename CT, int R>
struct CrossParams {
    typedef typename TabType<CT>::type TT;
    int64_t Nx, Nd, No;
    int tx, td, nbx, nbd;
    int mx, md;
    const CT* CX; CT* TY;
    int64_t nlx, nly;
    const TT* G;    // [V_COUNT][mx][TILE][R]   x response to a unit carry of scan q, after all later x scans
    const TT* L;    // [V_COUNT][md][R][TILE]   d tail of scan s per unit impulse at row i (scans 1..s)
};

} // namespace rfb
