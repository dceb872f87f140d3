/*
 * plan.cu -- launch planner and C ABI of the recursive-filter engine.
 *
 * Replaces lib/schedule.cpp + the GPU auto-schedules of lib/recfilter.cpp:682-870
 * (which map tagged Halide loop variables to CUDA blocks/threads) and the host
 * precompute of lib/coefficients.cpp:8-128 + lib/split.cpp:152-203 (tail weight
 * matrices).  The planner turns a scan list into passes, computes every carry
 * matrix by simulating the scans on unit vectors in fp64 (or in the integer ring
 * for integer filters), and drives the kernels of kernels.cu.
 *
 * There is no CPU execution path: without a CUDA device plan creation fails with
 * RF_ENODEVICE.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <new>
#include <string>
#include <vector>
#include "engine.h"
#include "fused_params.h"
#include "lookback_params.h"

namespace rfb {

// launchers defined in kernels.cu ---------------------------------------------------------------
#define RFB_FOR_EACH_R(X) X(1) X(2) X(3) X(4) X(8) X(16) X(32)
#define DECLARE_LAUNCHERS(RR)                                                                       \
    cudaError_t launch_tile_f##RR(const PassParams<float, RR>&, const void*, void*, int, cudaStream_t);   \
    cudaError_t launch_tile_u##RR(const PassParams<uint32_t, RR>&, const void*, void*, int, cudaStream_t);\
    cudaError_t launch_chain_f##RR(const ChainParams<float, RR>&, cudaStream_t);                    \
    cudaError_t launch_chain_u##RR(const ChainParams<uint32_t, RR>&, cudaStream_t);                 \
    cudaError_t launch_cross_f##RR(const CrossParams<float, RR>&, cudaStream_t);                    \
    cudaError_t launch_cross_u##RR(const CrossParams<uint32_t, RR>&, cudaStream_t);
RFB_FOR_EACH_R(DECLARE_LAUNCHERS)

#define RFB_FOR_EACH_FR(X) X(1) X(2) X(3) X(4) X(8)
#define DECLARE_FLAUNCHERS(RR)                                                                      \
    cudaError_t launch_fused_tile_f##RR(const FusedParams<float, RR>&, const void*, void*, int, int, cudaStream_t);    \
    cudaError_t launch_fused_tile_u##RR(const FusedParams<uint32_t, RR>&, const void*, void*, int, int, cudaStream_t); \
    cudaError_t launch_fused_stream_f##RR(const FStreamParams<float, RR>&, const void*, void*, cudaStream_t);          \
    cudaError_t launch_fchain_f##RR(const FChainParams<float, RR>&, cudaStream_t);                  \
    cudaError_t launch_fchain_u##RR(const FChainParams<uint32_t, RR>&, cudaStream_t);               \
    cudaError_t launch_fcross_f##RR(const FCrossParams<float, RR>&, int, cudaStream_t);             \
    cudaError_t launch_fcross_u##RR(const FCrossParams<uint32_t, RR>&, int, cudaStream_t);          \
    cudaError_t launch_flocal_f##RR(const FLocalParams<float, RR>&, cudaStream_t);                  \
    cudaError_t launch_flocal_u##RR(const FLocalParams<uint32_t, RR>&, cudaStream_t);
RFB_FOR_EACH_FR(DECLARE_FLAUNCHERS)

static inline cudaError_t launch_fused_stream(const FStreamParams<float, 1>& p, const void* i, void* o, cudaStream_t s) { return launch_fused_stream_f1(p, i, o, s); }
static inline cudaError_t launch_fused_stream(const FStreamParams<float, 2>& p, const void* i, void* o, cudaStream_t s) { return launch_fused_stream_f2(p, i, o, s); }
static inline cudaError_t launch_fused_stream(const FStreamParams<float, 3>& p, const void* i, void* o, cudaStream_t s) { return launch_fused_stream_f3(p, i, o, s); }
static inline cudaError_t launch_fused_stream(const FStreamParams<float, 4>& p, const void* i, void* o, cudaStream_t s) { return launch_fused_stream_f4(p, i, o, s); }

template <typename CT, int R> struct FLaunch;
#define DEFINE_FLAUNCH_TRAITS(RR)                                                                   \
    template <> struct FLaunch<float, RR> {                                                         \
        static cudaError_t tile(const FusedParams<float, RR>& p, const void* i, void* o, int m, int ts, cudaStream_t s) { return launch_fused_tile_f##RR(p, i, o, m, ts, s); } \
        static cudaError_t chain(const FChainParams<float, RR>& p, cudaStream_t s) { return launch_fchain_f##RR(p, s); } \
        static cudaError_t cross(const FCrossParams<float, RR>& p, int ts, cudaStream_t s) { return launch_fcross_f##RR(p, ts, s); } \
        static cudaError_t local(const FLocalParams<float, RR>& p, cudaStream_t s) { return launch_flocal_f##RR(p, s); } \
    };                                                                                              \
    template <> struct FLaunch<uint32_t, RR> {                                                      \
        static cudaError_t tile(const FusedParams<uint32_t, RR>& p, const void* i, void* o, int m, int ts, cudaStream_t s) { return launch_fused_tile_u##RR(p, i, o, m, ts, s); } \
        static cudaError_t chain(const FChainParams<uint32_t, RR>& p, cudaStream_t s) { return launch_fchain_u##RR(p, s); } \
        static cudaError_t cross(const FCrossParams<uint32_t, RR>& p, int ts, cudaStream_t s) { return launch_fcross_u##RR(p, ts, s); } \
        static cudaError_t local(const FLocalParams<uint32_t, RR>& p, cudaStream_t s) { return launch_flocal_u##RR(p, s); } \
    };
RFB_FOR_EACH_FR(DEFINE_FLAUNCH_TRAITS)

// single-pass look-back kernels (lookback_inst.cu)
#define DECLARE_LBLAUNCHERS(RR)                                                                     \
    cudaError_t launch_lb_tile_f##RR(const LBTileParams<float, RR>&, const void*, void*, int, cudaStream_t);      \
    cudaError_t launch_lb_tile_u##RR(const LBTileParams<uint32_t, RR>&, const void*, void*, int, cudaStream_t);   \
    cudaError_t launch_lb_signal_f##RR(const LBSignalParams<float, RR>&, const void*, void*, cudaStream_t);       \
    cudaError_t launch_lb_signal_u##RR(const LBSignalParams<uint32_t, RR>&, const void*, void*, cudaStream_t);
RFB_FOR_EACH_FR(DECLARE_LBLAUNCHERS)
template <typename CT, int R> struct LBLaunch;
#define DEFINE_LBLAUNCH_TRAITS(RR)                                                                  \
    template <> struct LBLaunch<float, RR> {                                                        \
        static cudaError_t tile(const LBTileParams<float, RR>& p, const void* i, void* o, int ts, cudaStream_t s) { return launch_lb_tile_f##RR(p, i, o, ts, s); } \
        static cudaError_t signal(const LBSignalParams<float, RR>& p, const void* i, void* o, cudaStream_t s) { return launch_lb_signal_f##RR(p, i, o, s); } \
    };                                                                                              \
    template <> struct LBLaunch<uint32_t, RR> {                                                     \
        static cudaError_t tile(const LBTileParams<uint32_t, RR>& p, const void* i, void* o, int ts, cudaStream_t s) { return launch_lb_tile_u##RR(p, i, o, ts, s); } \
        static cudaError_t signal(const LBSignalParams<uint32_t, RR>& p, const void* i, void* o, cudaStream_t s) { return launch_lb_signal_u##RR(p, i, o, s); } \
    };
RFB_FOR_EACH_FR(DEFINE_LBLAUNCH_TRAITS)

template <typename CT, int R> struct Launch;
#define DEFINE_LAUNCH_TRAITS(RR)                                                                    \
    template <> struct Launch<float, RR> {                                                          \
        static cudaError_t tile(const PassParams<float, RR>& p, const void* i, void* o, int m, cudaStream_t s) { return launch_tile_f##RR(p, i, o, m, s); } \
        static cudaError_t chain(const ChainParams<float, RR>& p, cudaStream_t s) { return launch_chain_f##RR(p, s); } \
        static cudaError_t cross(const CrossParams<float, RR>& p, cudaStream_t s) { return launch_cross_f##RR(p, s); } \
    };                                                                                              \
    template <> struct Launch<uint32_t, RR> {                                                       \
        static cudaError_t tile(const PassParams<uint32_t, RR>& p, const void* i, void* o, int m, cudaStream_t s) { return launch_tile_u##RR(p, i, o, m, s); } \
        static cudaError_t chain(const ChainParams<uint32_t, RR>& p, cudaStream_t s) { return launch_chain_u##RR(p, s); } \
        static cudaError_t cross(const CrossParams<uint32_t, RR>& p, cudaStream_t s) { return launch_cross_u##RR(p, s); } \
    };
RFB_FOR_EACH_R(DEFINE_LAUNCH_TRAITS)

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                   \
    do { cudaError_t e__ = (expr);                                                       \
         if (e__ != cudaSuccess)                                                         \
             return fail(RF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// TMA descriptors (fused path).  cuTensorMapEncodeTiled is a driver entry point; it is resolved
// through the runtime so that the library links against cudart only.
// ---------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t make_tile_map(CUtensorMap* map, const void* base, int64_t Nx, int64_t rows, int ts, bool is_float)
{
    static encode_tiled_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (!fn || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        encode = (encode_tiled_fn)fn;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return cudaErrorMisalignedAddress;
    const cuuint64_t dims[2]    = { (cuuint64_t)Nx, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)Nx * 4 };
    const cuuint32_t box[2]     = { 32u, (cuuint32_t)ts };
    const cuuint32_t estr[2]    = { 1u, 1u };
    const CUresult r = encode(map, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
                              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// host-side scan simulation (semantics identical to scan_regs in kernels.cu)
//   HT = double for float filters, uint32_t (wrapping ring) for integer filters
// ---------------------------------------------------------------------------------------------
//   scaled: the unit-feed-forward form of the fused path (scan_line in fused.cuh); c[0] is then
//   the clamp-history factor 1/b0 instead of b0
template <typename HT>
static void sim_scan(std::vector<HT>& v, int len, std::vector<HT>& h, const std::vector<HT>& c,
                     bool causal, bool clampb, bool scaled = false)
{
    const int R = (int)h.size();
    for (int p = 0; p < len; ++p) {
        const int i = causal ? p : len - 1 - p;
        const HT x = v[i];
        HT acc = scaled ? x : c[0] * x;
        if (p == 0 && clampb) {
            const HT tap = scaled ? x * c[0] : x;
            for (int k = 1; k <= R; ++k) acc = acc + c[k] * tap;
            for (int k = 0; k < R; ++k) h[k] = acc;
        } else {
            for (int k = 1; k <= R; ++k) acc = acc + c[k] * h[k - 1];
            for (int k = R - 1; k >= 1; --k) h[k] = h[k - 1];
            h[0] = acc;
        }
        v[i] = acc;
    }
}

struct HostScan {
    int causal;
    int order;
    float coeff[RF_MAX_ORDER + 1];
};

template <typename HT> static HT cvt_coeff(float c);
template <> double   cvt_coeff<double>(float c)   { return (double)c; }
template <> uint32_t cvt_coeff<uint32_t>(float c) { return (uint32_t)(int64_t)c; }   // Cast::make(type, coeff), wraps

template <typename HT> static HT clamp_factor(float b0);
template <> double   clamp_factor<double>(float b0)   { return (double)(float)(1.0 / (double)b0); }   // as the device holds it
template <> uint32_t clamp_factor<uint32_t>(float)    { return 1u; }                                  // b0 == 1 required

template <typename HT>
static std::vector<HT> coeff_vec(const HostScan& s, int R, bool scaled = false)
{
    std::vector<HT> c(R + 1, (HT)0);
    for (int k = 0; k <= s.order; ++k) c[k] = cvt_coeff<HT>(s.coeff[k]);
    if (scaled) c[0] = clamp_factor<HT>(s.coeff[0]);
    return c;
}

struct VariantGeom { int len; int lo; int hi; };   // lo/hi: 1 = closed image border on that face

static VariantGeom variant_geom(const DimGeom& g, int var)
{
    switch (var) {
    case V_FIRST:    return { g.t, g.lo_closed, 0 };
    case V_INTERIOR: return { g.t, 0, 0 };
    case V_LAST:     return { g.len_last, 0, g.hi_closed };
    default:         return { g.len_last, g.lo_closed, g.hi_closed };   // V_SINGLE
    }
}

// carry matrices of one dimension
template <typename HT>
struct DimTables {
    int S = 0, R = 0;
    std::vector<HT> P;      // [V][S][R][R]
    std::vector<HT> M;      // [V][S][S][R][R]
    std::vector<HT> G;      // [V][S][TILE][R]
    std::vector<HT> L;      // [V][S][R][TILE]
    std::vector<HT> Pseg;   // [S][2][R][R]
};

template <typename HT>
static void matmul_rr(std::vector<HT>& out, const HT* a, const HT* b, int R)   // out = a * b
{
    out.assign((size_t)R * R, (HT)0);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < R; ++j) {
            HT acc = (HT)0;
            for (int k = 0; k < R; ++k) acc = acc + a[i * R + k] * b[k * R + j];
            out[i * R + j] = acc;
        }
}

template <typename HT>
static void build_dim_tables(DimTables<HT>& tb, const std::vector<HostScan>& scans, const DimGeom& g,
                             int R, bool clamp, int seg, int nseg, bool want_GL, bool scaled = false, int TILE = rfb::TILE)
{
    const int S = (int)scans.size();
    tb.S = S; tb.R = R;
    tb.P.assign((size_t)V_COUNT * S * R * R, (HT)0);
    tb.M.assign((size_t)V_COUNT * S * S * R * R, (HT)0);
    if (want_GL) {
        tb.G.assign((size_t)V_COUNT * S * TILE * R, (HT)0);
        tb.L.assign((size_t)V_COUNT * S * R * TILE, (HT)0);
    }
    std::vector<std::vector<HT>> coef(S);
    for (int s = 0; s < S; ++s) coef[s] = coeff_vec<HT>(scans[s], R, scaled);

    auto closed = [&](const VariantGeom& vg, int s) { return scans[s].causal ? vg.lo : vg.hi; };

    for (int var = 0; var < V_COUNT; ++var) {
        const VariantGeom vg = variant_geom(g, var);
        const int len = vg.len;
        if (len <= 0) continue;
        for (int q = 0; q < S; ++q) {
            if (closed(vg, q)) continue;            // no carry enters this tile for scan q
            for (int kk = 0; kk < R; ++kk) {
                std::vector<HT> v(len, (HT)0), h(R, (HT)0);
                h[kk] = (HT)1;
                sim_scan<HT>(v, len, h, coef[q], scans[q].causal != 0, false, scaled);
                for (int k = 0; k < R; ++k) tb.P[(((size_t)var * S + q) * R + k) * R + kk] = h[k];
                // propagate the response through the later scans of this dimension
                for (int s = q + 1; s < S; ++s) {
                    std::vector<HT> hs(R, (HT)0);
                    sim_scan<HT>(v, len, hs, coef[s], scans[s].causal != 0, clamp && closed(vg, s), scaled);
                    for (int k = 0; k < R; ++k)
                        tb.M[((((size_t)var * S + q) * S + s) * R + k) * R + kk] = hs[k];
                }
                if (want_GL && len <= TILE)
                    for (int i = 0; i < len; ++i) tb.G[(((size_t)var * S + q) * TILE + i) * R + kk] = v[i];
            }
        }
        if (want_GL && len <= TILE) {
            for (int i = 0; i < len; ++i) {
                std::vector<HT> v(len, (HT)0);
                v[i] = (HT)1;
                for (int s = 0; s < S; ++s) {
                    std::vector<HT> hs(R, (HT)0);
                    sim_scan<HT>(v, len, hs, coef[s], scans[s].causal != 0, clamp && closed(vg, s), scaled);
                    for (int k = 0; k < R; ++k) tb.L[(((size_t)var * S + s) * R + k) * TILE + i] = hs[k];
                }
            }
        }
    }

    // segment products for the two-level chain
    tb.Pseg.assign((size_t)S * 2 * R * R, (HT)0);
    for (int s = 0; s < S && seg > 0; ++s) {
        const HT* Pint = &tb.P[(((size_t)V_INTERIOR * S + s)) * R * R];
        std::vector<HT> acc((size_t)R * R, (HT)0), tmp;
        for (int k = 0; k < R; ++k) acc[k * R + k] = (HT)1;
        for (int i = 0; i < seg; ++i) { matmul_rr<HT>(tmp, Pint, acc.data(), R); acc = tmp; }
        std::copy(acc.begin(), acc.end(), tb.Pseg.begin() + ((size_t)s * 2 + 0) * R * R);
        // last segment in scan order
        std::fill(acc.begin(), acc.end(), (HT)0);
        for (int k = 0; k < R; ++k) acc[k * R + k] = (HT)1;
        for (int jj = (nseg - 1) * seg; jj < g.nb; ++jj) {
            const int j = scans[s].causal ? jj : g.nb - 1 - jj;
            const int var = g.nb == 1 ? V_SINGLE : (j == 0 ? V_FIRST : (j == g.nb - 1 ? V_LAST : V_INTERIOR));
            matmul_rr<HT>(tmp, &tb.P[((size_t)var * S + s) * R * R], acc.data(), R);
            acc = tmp;
        }
        std::copy(acc.begin(), acc.end(), tb.Pseg.begin() + ((size_t)s * 2 + 1) * R * R);
    }

}

// whole-dimension matrices of one shard: response of the shard's outgoing tails to its
// incoming carries.  Pdim [S][R][R], Mdim [S][S][R][R].
template <typename HT>
static void build_whole_dim(std::vector<HT>& Pdim, std::vector<HT>& Mdim, const std::vector<HostScan>& scans,
                            int64_t n, int lo_closed, int hi_closed, int R, bool clamp, bool scaled = false)
{
    const int S = (int)scans.size();
    Pdim.assign((size_t)S * R * R, (HT)0);
    Mdim.assign((size_t)S * S * R * R, (HT)0);
    std::vector<std::vector<HT>> coef(S);
    for (int s = 0; s < S; ++s) coef[s] = coeff_vec<HT>(scans[s], R, scaled);
    const int len = (int)n;
    auto dclosed = [&](int s) { return scans[s].causal ? lo_closed : hi_closed; };
    for (int q = 0; q < S; ++q) {
        if (dclosed(q)) continue;
        for (int kk = 0; kk < R; ++kk) {
            std::vector<HT> v(len, (HT)0), h(R, (HT)0);
            h[kk] = (HT)1;
            sim_scan<HT>(v, len, h, coef[q], scans[q].causal != 0, false, scaled);
            for (int k = 0; k < R; ++k) Pdim[((size_t)q * R + k) * R + kk] = h[k];
            for (int s = q + 1; s < S; ++s) {
                std::vector<HT> hs(R, (HT)0);
                sim_scan<HT>(v, len, hs, coef[s], scans[s].causal != 0, clamp && dclosed(s), scaled);
                for (int k = 0; k < R; ++k) Mdim[(((size_t)q * S + s) * R + k) * R + kk] = hs[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { bytes = n; if (n == 0) return cudaSuccess; return cudaMalloc(&p, n); }
    DevBuf() = default; DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
};

template <typename HT, typename CT>
static cudaError_t upload(DevBuf& b, const std::vector<HT>& src)
{
    std::vector<CT> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (CT)src[i];
    cudaError_t e = b.alloc(tmp.size() * sizeof(CT));
    if (e != cudaSuccess || tmp.empty()) return e;
    return cudaMemcpy(b.p, tmp.data(), tmp.size() * sizeof(CT), cudaMemcpyHostToDevice);
}

// optional per-stage device timing (CUDA events on the launching stream)
enum { ST_TAILS = 0, ST_CHAIN = 1, ST_CROSS = 2, ST_FINAL = 3, ST_CONVERT = 4, ST_COUNT = 5 };
struct StageTimer {
    bool on = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    double ms[ST_COUNT] = { 0, 0, 0, 0, 0 };
    long   count[ST_COUNT] = { 0, 0, 0, 0, 0 };
    cudaEvent_t begin(cudaStream_t st, int stage)
    {
        if (!on) return nullptr;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        pending.push_back({ stage, { a, b } });
        return b;
    }
    void end(cudaStream_t st, cudaEvent_t b) { if (b) cudaEventRecord(b, st); }
    void collect()
    {
        for (auto& e : pending) {
            cudaEventSynchronize(e.second.second);
            float t = 0.f;
            cudaEventElapsedTime(&t, e.second.first, e.second.second);
            ms[e.first] += t; count[e.first] += 1;
            cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second);
        }
        pending.clear();
    }
    void reset() { collect(); for (int i = 0; i < ST_COUNT; ++i) { ms[i] = 0; count[i] = 0; } }
};

// ---------------------------------------------------------------------------------------------
// strip sharding without re-running the carry chain.  Stage 1 chains a strip with ZERO carries entering its
// open faces and keeps those carries C0.  The carries are linear in what enters the strip:
//     C_s[j] = C0_s[j] + sum_{q <= s} W[s][q][j] * ext_q          (j: tile of the strip, R x R blocks)
// W is simulated once at plan creation (unit histories entering the strip, fp64, conjugated to the difference
// basis of the fused carry algebra); after the exchange one elementwise kernel adds the second term.
// ---------------------------------------------------------------------------------------------
template <typename HT>
static void build_strip_response(std::vector<HT>& W, const std::vector<HostScan>& scans, int64_t n, int ts, int nb,
                                 int lo_closed, int hi_closed, int R, bool clamp, bool scaled)
{
    const int S = (int)scans.size();
    W.assign((size_t)S * S * nb * R * R, (HT)0);
    std::vector<std::vector<HT>> coef(S);
    for (int s = 0; s < S; ++s) coef[s] = coeff_vec<HT>(scans[s], R, scaled);
    const int len = (int)n;
    auto closed = [&](int s) { return scans[s].causal ? lo_closed : hi_closed; };
    // history entering tile j of scan s, read off the array after scan s (k = 0 newest)
    auto entering = [&](const std::vector<HT>& v, int s, int j, int k) -> HT {
        if (scans[s].causal) { const int i = j * ts - 1 - k; return i >= 0 ? v[i] : (HT)0; }
        const int i = (j + 1) * ts + k; return i < len ? v[i] : (HT)0;
    };
    for (int q = 0; q < S; ++q) {
        if (closed(q)) continue;                     // nothing enters the strip for this scan
        for (int kk = 0; kk < R; ++kk) {
            std::vector<HT> v(len, (HT)0), h(R, (HT)0);
            h[kk] = (HT)1;
            sim_scan<HT>(v, len, h, coef[q], scans[q].causal != 0, false, scaled);
            for (int j = 0; j < nb; ++j) {
                const bool entry = scans[q].causal ? (j == 0) : (j == nb - 1);
                for (int k = 0; k < R; ++k)
                    W[((((size_t)q * S + q) * nb + j) * R + k) * R + kk] = entry ? (HT)(k == kk ? 1 : 0) : entering(v, q, j, k);
            }
            for (int s = q + 1; s < S; ++s) {
                std::vector<HT> hs(R, (HT)0);
                sim_scan<HT>(v, len, hs, coef[s], scans[s].causal != 0, clamp && closed(s), scaled);
                for (int j = 0; j < nb; ++j)
                    for (int k = 0; k < R; ++k)
                        W[((((size_t)s * S + q) * nb + j) * R + k) * R + kk] = entering(v, s, j, k);
            }
        }
    }
}

// CY[s][k][j][l] += D^-1 ( sum_{q <= s} W'[s][q][j] * D ext_q[.][l] ): one thread per (line, tile).  active[s * nb + j]
// is 0 where every entry of W[s][.][j] rounds to zero in the compute type (a short-memory filter only reaches the
// first few tiles of the strip: those blocks are skipped, which changes nothing).
template <typename CT, int R>
__global__ void carry_fix_kernel(CT* __restrict__ CY, const CT* __restrict__ ext, const CT* __restrict__ Wd,
                                 const unsigned char* __restrict__ active, int S, int nb, int64_t nl)
{
    const int j = blockIdx.y;
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool any = false;
    for (int s = 0; s < S; ++s) any = any || active[s * nb + j];
    if (!any || l >= nl) return;
    CT e[FMAX_SCANS][R];
    for (int q = 0; q < S; ++q) {
        CT h[R];
#pragma unroll
        for (int k = 0; k < R; ++k) h[k] = ext[((int64_t)q * R + k) * nl + l];
#pragma unroll
        for (int m = 1; m < R; ++m)
#pragma unroll
            for (int k = R - 1; k >= m; --k) h[k] = h[k - 1] - h[k];               // difference basis
#pragma unroll
        for (int k = 0; k < R; ++k) e[q][k] = h[k];
    }
    for (int s = 0; s < S; ++s) {
        if (!active[s * nb + j]) continue;
        CT dc[R];
#pragma unroll
        for (int k = 0; k < R; ++k) dc[k] = (CT)0;
        for (int q = 0; q <= s; ++q) {
            const CT* w = Wd + (((size_t)s * S + q) * nb + j) * R * R;
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int kk = 0; kk < R; ++kk) dc[k] = dc[k] + __ldg(w + k * R + kk) * e[q][kk];
        }
#pragma unroll
        for (int m = R - 1; m >= 1; --m)
#pragma unroll
            for (int k = m; k < R; ++k) dc[k] = dc[k - 1] - dc[k];                 // back to histories
#pragma unroll
        for (int k = 0; k < R; ++k) {
            CT* c = CY + (((int64_t)s * R + k) * nb + j) * nl + l;
            *c = *c + dc[k];
        }
    }
}

// small kernel: strip-level carry resolution (host of the multi-GPU layer)
template <typename CT, int R>
__global__ void shard_resolve_kernel(const CT* __restrict__ tails, CT* __restrict__ ext, int64_t nl, int S,
                                     int nshards, int rank, const typename TabType<CT>::type* __restrict__ Pdim3,
                                     const typename TabType<CT>::type* __restrict__ Mdim3,
                                     const int* __restrict__ causal)
{
    typedef typename TabType<CT>::type TT;
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    // c[q][g][k] kept in local memory (S and nshards are tiny)
    TT c[8][8][R];
    const int64_t shard_stride = (int64_t)S * R * nl;
    for (int s = 0; s < S; ++s) {
        TT tau[R];
        for (int k = 0; k < R; ++k) tau[k] = (TT)0;
        for (int gg = 0; gg < nshards; ++gg) {
            const int g = causal[s] ? gg : nshards - 1 - gg;
            const int pos = (g == 0) ? 0 : (g == nshards - 1 ? 2 : 1);     // first / interior / last shard
            const TT* Pdim = Pdim3 + (int64_t)pos * S * R * R;
            const TT* Mdim = Mdim3 + (int64_t)pos * S * S * R * R;
            for (int k = 0; k < R; ++k) c[s][g][k] = tau[k];
            TT t[R];
            for (int k = 0; k < R; ++k) t[k] = (TT)tails[g * shard_stride + ((int64_t)s * R + k) * nl + l];
            for (int q = 0; q < s; ++q)
                for (int k = 0; k < R; ++k)
                    for (int kk = 0; kk < R; ++kk)
                        t[k] = t[k] + Mdim[(((int64_t)q * S + s) * R + k) * R + kk] * c[q][g][kk];
            for (int k = 0; k < R; ++k)
                for (int kk = 0; kk < R; ++kk)
                    t[k] = t[k] + Pdim[((int64_t)s * R + k) * R + kk] * tau[kk];
            for (int k = 0; k < R; ++k) tau[k] = t[k];
        }
        if (rank >= 0) {
            for (int k = 0; k < R; ++k) ext[((int64_t)s * R + k) * nl + l] = (CT)c[s][rank][k];
        } else {                                   // rank < 0: the carries entering EVERY shard, ext[g][s][k][l]
            for (int gg = 0; gg < nshards; ++gg)
                for (int k = 0; k < R; ++k) ext[gg * shard_stride + ((int64_t)s * R + k) * nl + l] = (CT)c[s][gg][k];
        }
    }
}

// Every shard holds the zero-history tails of all shards: tails[g][s][k][ly].  Resolve the carries
// entering this shard with the whole-dimension matrices of a first / interior / last shard
// (built once, at plan creation).
template <typename CT, int R>
struct ShardResolver {
    using HT = typename std::conditional<std::is_same<CT, float>::value, double, uint32_t>::type;
    using TT = typename TabType<CT>::type;
    DevBuf dP, dM, dC;
    int md = 0;
    bool ready = false;
    // every entry of the whole-shard transition matrices is below 1e-30 (fp64): what enters a shard is decided by the
    // tails of the two adjacent shards alone, farther shards contribute P * (...) = nothing an fp32 carry can hold
    bool neighbors_suffice = false;
    int init(const std::vector<HostScan>& sd, int64_t n, bool clamp, bool scaled)
    {
        md = (int)sd.size();
        if (md > 8) return fail(RF_EUNSUPPORTED, "sharded execution supports <= 8 scans along the sharded dimension");
        std::vector<HT> P3, M3, Pt, Mt;
        for (int pos = 0; pos < 3; ++pos) {
            build_whole_dim<HT>(Pt, Mt, sd, n, pos == 0 ? 1 : 0, pos == 2 ? 1 : 0, R, clamp, scaled);
            P3.insert(P3.end(), Pt.begin(), Pt.end());
            M3.insert(M3.end(), Mt.begin(), Mt.end());
        }
        neighbors_suffice = std::is_same<CT, float>::value;
        for (const HT& v : P3) if (!(std::fabs((double)v) < 1e-30)) neighbors_suffice = false;
        CUDA_TRY((upload<HT, TT>(dP, P3)));
        CUDA_TRY((upload<HT, TT>(dM, M3)));
        std::vector<int> causal(md);
        for (int s = 0; s < md; ++s) causal[s] = sd[s].causal;
        CUDA_TRY(dC.alloc(causal.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(dC.p, causal.data(), causal.size() * sizeof(int), cudaMemcpyHostToDevice));
        ready = true;
        return RF_OK;
    }
    // rank < 0: resolve the carries entering every shard for the given lines (ext_out is [nshards][S][R][nl])
    int run(int64_t nl, const void* gathered, int nshards, int rank, void* ext_out, cudaStream_t st)
    {
        if (!ready) return fail(RF_EINVAL, "plan was not created for sharded execution");
        if (nshards > 8) return fail(RF_EUNSUPPORTED, "sharded execution supports <= 8 shards");
        shard_resolve_kernel<CT, R><<<(unsigned)((nl + 127) / 128), 128, 0, st>>>(
            (const CT*)gathered, (CT*)ext_out, nl, md, nshards, rank, (const TT*)dP.p, (const TT*)dM.p,
            (const int*)dC.p);
        CUDA_TRY(cudaGetLastError());
        return RF_OK;
    }
};

struct PassBase {
    StageTimer* timer = nullptr;
    virtual ~PassBase() {}
    virtual int run_tails(const void* in, void* out, cudaStream_t st) = 0;            // K1
    virtual int run_carries(const void* ext_d, void* tail_out_d, cudaStream_t st, int stage = 0) = 0; // K2/K3 (ext: shard carries; stage 1/2 of a sharded run)
    virtual int run_final(const void* in, void* out, cudaStream_t st) = 0;            // K4
    virtual bool needs_carries() const = 0;
    virtual int launches() const = 0;
    virtual size_t workspace() const = 0;
    virtual std::string describe() const = 0;
    // sharding support (column dimension only)
    virtual bool d_open() const = 0;
    virtual size_t shard_tail_elems() const = 0;     // elements of the compute type exported per shard
    virtual int shard_resolve(const void* gathered, int nshards, int rank, cudaStream_t st) = 0;
    // carries entering every shard for `nlines` lines of the cut (column-chunked exchange); vectors = scans x order
    virtual int shard_resolve_lines(const void* gathered, int nshards, int64_t nlines, void* ext_all, cudaStream_t st) = 0;
    virtual int shard_vectors() const = 0;
    virtual bool shard_neighbors_suffice() const { return false; }
    virtual const void* ext_buffer() const = 0;
    // device word a kernel of the pass sets when it gave up waiting (look-back kernels), or null
    virtual const uint32_t* error_flag() const { return nullptr; }
    // pointwise epilogue out = a_out * filtered + a_in * input fused into the store of the last kernel (rf_options.epilogue)
    virtual bool set_epilogue(float, float) { return false; }
    // the whole pass on one stream; a pass may overlap its own stages internally (FusedPass: stack slices)
    virtual int run_all(const void* in, void* out, cudaStream_t st)
    {
        int rc;
        if ((rc = run_tails(in, out, st))) return rc;
        if ((rc = run_carries(nullptr, nullptr, st))) return rc;
        return run_final(in, out, st);
    }
};

template <typename CT, int R>
struct Pass : PassBase {
    using HT = typename std::conditional<std::is_same<CT, float>::value, double, uint32_t>::type;
    using TT = typename TabType<CT>::type;     // device table / accumulation type (== HT)
    int dtype = RF_F32;
    PassParams<CT, R> pp;
    DimGeom gx, gd;
    std::vector<HostScan> sx, sd;
    bool fused = false;
    int segx = 1, nsegx = 1, segd = 1, nsegd = 1;
    DevBuf TX, CX, TY, CY, SEGT, SEGC;
    DevBuf dPx, dMx, dPsegx, dPd, dMd, dPsegd, dG, dL;
    DevBuf dExt, dTailOut;           // shard carries in / tails out  [s][k][ly]
    DimTables<HT> tx_tab, td_tab;
    std::unique_ptr<ShardResolver<CT, R>> resolver;

    bool x_needs() const { return gx.nscans > 0 && gx.nb > 1; }
    bool d_needs() const { return gd.nscans > 0 && (gd.nb > 1 || !gd.lo_closed || !gd.hi_closed); }
    bool needs_carries() const override { return x_needs() || d_needs(); }
    bool d_open() const override { return gd.nscans > 0 && (!gd.lo_closed || !gd.hi_closed); }
    const void* ext_buffer() const override { return dExt.p; }

    size_t workspace() const override
    {
        return TX.bytes + CX.bytes + TY.bytes + CY.bytes + SEGT.bytes + SEGC.bytes + dPx.bytes + dMx.bytes +
               dPsegx.bytes + dPd.bytes + dMd.bytes + dPsegd.bytes + dG.bytes + dL.bytes + dExt.bytes + dTailOut.bytes;
    }

    int launches() const override
    {
        int n = 1;                                   // K4
        if (needs_carries()) {
            n += 1;                                  // K1
            if (x_needs()) n += gx.nscans * (nsegx > 1 ? 3 : 1);
            if (x_needs() && d_needs() && fused) n += 1;
            if (d_needs()) n += gd.nscans * (nsegd > 1 ? 3 : 1);
        }
        return n;
    }

    int init(int64_t Nx, int64_t Nd, int64_t No, bool signal, bool clamp)
    {
        // geometry -> params
        std::memset(&pp, 0, sizeof(pp));
        pp.Nx = Nx; pp.Nd = Nd; pp.No = No; pp.nlx = Nd * No; pp.nly = Nx * No;
        pp.signal_mode = signal ? 1 : 0;
        pp.tx = gx.t; pp.td = gd.t; pp.nbx = gx.nb; pp.nbd = gd.nb;
        pp.lenx_last = gx.len_last; pp.lend_last = gd.len_last;
        pp.clamp = clamp ? 1 : 0;
        pp.x_lo_closed = gx.lo_closed; pp.x_hi_closed = gx.hi_closed;
        pp.d_lo_closed = gd.lo_closed; pp.d_hi_closed = gd.hi_closed;
        pp.mx = gx.nscans; pp.md = gd.nscans;
        for (int s = 0; s < pp.mx; ++s) {
            pp.sx.causal[s] = sx[s].causal;
            for (int k = 0; k <= R; ++k) pp.sx.coef[s][k] = k <= sx[s].order ? (CT)cvt_coeff<HT>(sx[s].coeff[k]) : (CT)0;
        }
        for (int s = 0; s < pp.md; ++s) {
            pp.sd.causal[s] = sd[s].causal;
            for (int k = 0; k <= R; ++k) pp.sd.coef[s][k] = k <= sd[s].order ? (CT)cvt_coeff<HT>(sd[s].coeff[k]) : (CT)0;
        }
        auto pick_seg = [](int nb, int& seg, int& nseg) {
            if (nb <= 512) { seg = nb; nseg = 1; }
            else { seg = 256; nseg = (nb + seg - 1) / seg; }
        };
        pick_seg(gx.nb, segx, nsegx);
        pick_seg(gd.nb, segd, nsegd);

        if (pp.mx > 0) {
            build_dim_tables<HT>(tx_tab, sx, gx, R, clamp, segx, nsegx, fused);
            CUDA_TRY((upload<HT, TT>(dPx, tx_tab.P)));
            CUDA_TRY((upload<HT, TT>(dMx, tx_tab.M)));
            CUDA_TRY((upload<HT, TT>(dPsegx, tx_tab.Pseg)));
            if (fused) CUDA_TRY((upload<HT, TT>(dG, tx_tab.G)));
            const size_t n = (size_t)pp.mx * R * gx.nb * pp.nlx * sizeof(CT);
            CUDA_TRY(TX.alloc(n)); CUDA_TRY(CX.alloc(n));
            CUDA_TRY(cudaMemset(CX.p, 0, n));
        }
        if (pp.md > 0) {
            build_dim_tables<HT>(td_tab, sd, gd, R, clamp, segd, nsegd, fused);
            CUDA_TRY((upload<HT, TT>(dPd, td_tab.P)));
            CUDA_TRY((upload<HT, TT>(dMd, td_tab.M)));
            CUDA_TRY((upload<HT, TT>(dPsegd, td_tab.Pseg)));
            if (fused) CUDA_TRY((upload<HT, TT>(dL, td_tab.L)));
            const size_t n = (size_t)pp.md * R * gd.nb * pp.nly * sizeof(CT);
            CUDA_TRY(TY.alloc(n)); CUDA_TRY(CY.alloc(n));
            CUDA_TRY(cudaMemset(CY.p, 0, n));
            if (d_open()) {
                const size_t m = (size_t)pp.md * R * pp.nly * sizeof(CT);
                CUDA_TRY(dExt.alloc(m)); CUDA_TRY(dTailOut.alloc(m));
                CUDA_TRY(cudaMemset(dExt.p, 0, m));
                resolver.reset(new ShardResolver<CT, R>());
                int rc = resolver->init(sd, gd.n, clamp, false);
                if (rc) return rc;
            }
        }
        {
            const size_t a = (size_t)R * nsegx * pp.nlx, b = (size_t)R * nsegd * pp.nly;
            const size_t n = std::max(nsegx > 1 ? a : 0, nsegd > 1 ? b : 0) * sizeof(TT);
            CUDA_TRY(SEGT.alloc(n)); CUDA_TRY(SEGC.alloc(n));
        }
        pp.TX = (CT*)TX.p; pp.CX = (CT*)CX.p; pp.TY = (CT*)TY.p; pp.CY = (CT*)CY.p;
        return RF_OK;
    }

    int run_tails(const void* in, void* out, cudaStream_t st) override
    {
        if (!needs_carries()) return RF_OK;
        cudaEvent_t ev = timer ? timer->begin(st, ST_TAILS) : nullptr;
        CUDA_TRY((Launch<CT, R>::tile(pp, in, out, 0, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    int run_chain(bool xdim, const void* ext_d, void* tail_out_d, cudaStream_t st)
    {
        const DimGeom& g = xdim ? gx : gd;
        ChainParams<CT, R> cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.T = (CT*)(xdim ? TX.p : TY.p);
        cp.C = (CT*)(xdim ? CX.p : CY.p);
        cp.nl = xdim ? pp.nlx : pp.nly;
        cp.nb = g.nb;
        if (xdim && pp.signal_mode) { cp.tile_stride = 1; cp.line_stride = g.nb; }
        else                        { cp.tile_stride = cp.nl; cp.line_stride = 1; }
        cp.plane = (int64_t)g.nb * cp.nl;
        cp.seg = xdim ? segx : segd; cp.nseg = xdim ? nsegx : nsegd;
        cp.P = (const TT*)(xdim ? dPx.p : dPd.p);
        cp.M = (const TT*)(xdim ? dMx.p : dMd.p);
        cp.S = g.nscans;
        cp.SEGT = (TT*)SEGT.p; cp.SEGC = (TT*)SEGC.p;
        const auto& scans = xdim ? sx : sd;
        for (int s = 0; s < g.nscans; ++s) {
            cp.s = s;
            cp.causal = scans[s].causal;
            cp.Pseg = (const TT*)(xdim ? dPsegx.p : dPsegd.p) + (size_t)s * 2 * R * R;
            cp.ext = (!xdim && ext_d) ? (const CT*)ext_d + (size_t)s * R * cp.nl : nullptr;
            cp.tail_out = (!xdim && tail_out_d) ? (CT*)tail_out_d + (size_t)s * R * cp.nl : nullptr;
            cudaEvent_t ev = timer ? timer->begin(st, ST_CHAIN) : nullptr;
            CUDA_TRY((Launch<CT, R>::chain(cp, st)));
            if (timer) timer->end(st, ev);
        }
        return RF_OK;
    }

    int run_carries(const void* ext_d, void* tail_out_d, cudaStream_t st, int stage) override
    {
        if (!needs_carries()) return RF_OK;
        // stage 2 of a sharded run: the x carries and the cross residual (added to TY in place) are
        // already complete from stage 1; only the d chain is redone with the incoming shard carries
        if (stage == 2) return d_needs() ? run_chain(false, ext_d, tail_out_d, st) : RF_OK;
        if (x_needs()) { int rc = run_chain(true, nullptr, nullptr, st); if (rc) return rc; }
        if (x_needs() && d_needs() && fused) {
            CrossParams<CT, R> cr;
            std::memset(&cr, 0, sizeof(cr));
            cr.Nx = pp.Nx; cr.Nd = pp.Nd; cr.No = pp.No;
            cr.tx = pp.tx; cr.td = pp.td; cr.nbx = pp.nbx; cr.nbd = pp.nbd;
            cr.mx = pp.mx; cr.md = pp.md;
            cr.CX = (const CT*)CX.p; cr.TY = (CT*)TY.p;
            cr.nlx = pp.nlx; cr.nly = pp.nly;
            cr.G = (const TT*)dG.p; cr.L = (const TT*)dL.p;
            cudaEvent_t ev = timer ? timer->begin(st, ST_CROSS) : nullptr;
            CUDA_TRY((Launch<CT, R>::cross(cr, st)));
            if (timer) timer->end(st, ev);
        }
        if (d_needs()) { int rc = run_chain(false, ext_d, tail_out_d, st); if (rc) return rc; }
        return RF_OK;
    }

    int run_final(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = timer ? timer->begin(st, ST_FINAL) : nullptr;
        CUDA_TRY((Launch<CT, R>::tile(pp, in, out, 1, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    size_t shard_tail_elems() const override { return d_open() ? (size_t)pp.md * R * pp.nly : 0; }

    // Every shard holds the zero-history tails of all shards: tails[g][s][k][ly].
    // Resolve this shard's incoming carries with the whole-dimension matrices.
    int shard_resolve(const void* gathered, int nshards, int rank, cudaStream_t st) override
    {
        if (!d_open()) return RF_OK;
        return resolver->run(pp.nly, gathered, nshards, rank, dExt.p, st);
    }
    int shard_resolve_lines(const void* gathered, int nshards, int64_t nlines, void* ext_all, cudaStream_t st) override
    {
        if (!d_open()) return RF_OK;
        return resolver->run(nlines, gathered, nshards, -1, ext_all, st);
    }
    int shard_vectors() const override { return pp.md * R; }
    bool shard_neighbors_suffice() const override { return resolver && resolver->neighbors_suffice; }

    std::string describe() const override
    {
        char b[512];
        snprintf(b, sizeof(b),
                 "  pass view [%lld][%lld][%lld] %s%s: x scans %d (tile %d, %d tiles), d scans %d (tile %d, %d tiles), "
                 "order<=%d, launches %d\n",
                 (long long)pp.No, (long long)pp.Nd, (long long)pp.Nx, fused ? "fused " : "",
                 pp.signal_mode ? "signal-mode" : "image-mode", pp.mx, pp.tx, pp.nbx, pp.md, pp.td, pp.nbd, R,
                 launches());
        return b;
    }
};

// tiles per chain thread: long segments cost fewer instructions per line, short ones expose more
// threads (and need less shared memory when a dimension has many scans)
static int fchain_pick_L(int nb, int S)
{
    if (const char* e = getenv("RFB_CHAIN_L")) { const int v = atoi(e); if ((v == 4 || v == FCHAIN_L) && nb <= 16 * v) return v; }
    return (nb <= 32 || (S > 2 && nb <= 64)) ? 4 : FCHAIN_L;
}
// does the carry chain of a fused pass fit the shared memory of one CTA?
static bool fchain_fits(int R, int Sx, int Sd, int nbx, int nbd)
{
    const size_t limit = 227u * 1024u;
    if (Sd > 0) {
        const int L = fchain_pick_L(nbd, Sd);
        if (fchain_smem_bytes(Sd, (nbd + L - 1) / L, R, L, nbd, 0) > limit) return false;
    }
    if (Sx > 0) {
        const int L = fchain_pick_L(nbx, Sx);
        const int sdk = Sd > 0 ? ((Sd * R + 3) / 4) * 4 : 0;
        if (fchain_smem_bytes(Sx, (nbx + L - 1) / L, R, L, nbx, sdk) > limit) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// the fused fast path (kernels in fused.cuh): full TS x TS tiles, unit-feed-forward scans,
// d scans before x scans, one carry-chain launch per dimension
// ---------------------------------------------------------------------------------------------
template <typename CT, int R>
struct FusedPass : PassBase {
    using HT = typename std::conditional<std::is_same<CT, float>::value, double, uint32_t>::type;
    using TT = CT;                  // fp32 / u32 carry algebra, tables rounded once from the fp64 host build
    FusedParams<CT, R> fp;
    int ts = 128;
    DimGeom gx, gd;
    std::vector<HostScan> sx, sd;
    bool clamp = false;
    int nsegx = 1, nsegd = 1;
    int Lx = FCHAIN_L, Ld = FCHAIN_L;            // tiles per chain thread
    bool local_x = false, local_d = false;       // short-memory dimension: carries from the adjacent tile only (flocal_kernel)
    bool local_p2 = false;                       // ... in every scanned dimension: pass 2 derives its carries itself, no carry kernels
    bool local_now = false;                      // the carry stage of the call in flight took that path
    int local_mode = 1;                          // RFB_LOCAL_P2: 1 = pass 2 applies the cross residual, 2 = the cross kernel corrects the x tails
    DevBuf TX, CX, TY, CY, dA;
    DevBuf dPx, dMx, dPsegx, dL, dPd, dMd, dPsegd, dG;
    DevBuf dExt, dTailOut, dW, dWact;       // dW: response of the strip's carries to what enters it (build_strip_response)
    // both sweeps in one launch (fused_stream_kernel): short memory in both dimensions, full 128-sample tiles, unsharded
    bool stream_ok = false;
    int lag_a = 3, lag_p = 6;
    DevBuf dSync, dStreamErr;               // ticket + per-row counters (zeroed before every launch); the give-up flag
    DimTables<HT> tx_tab, td_tab;
    std::unique_ptr<ShardResolver<CT, R>> resolver;
    int sdk() const { return ((fp.md * R + 3) / 4) * 4; }     // entries of one A row, padded for 128-bit loads

    bool x_needs() const { return gx.nscans > 0 && gx.nb > 1; }
    bool d_needs() const { return gd.nscans > 0 && (gd.nb > 1 || !gd.lo_closed || !gd.hi_closed); }
    bool needs_carries() const override { return x_needs() || d_needs(); }
    bool d_open() const override { return gd.nscans > 0 && (!gd.lo_closed || !gd.hi_closed); }
    const void* ext_buffer() const override { return dExt.p; }
    bool cross_needed() const { return x_needs() && d_needs(); }

    size_t workspace() const override
    {
        return TX.bytes + CX.bytes + TY.bytes + CY.bytes + dA.bytes + dPx.bytes + dMx.bytes + dPsegx.bytes + dL.bytes +
               dPd.bytes + dMd.bytes + dPsegd.bytes + dG.bytes + dExt.bytes + dTailOut.bytes + dW.bytes + dWact.bytes;
    }
    int launches() const override
    {
        if (stream_ok && nslices < 2) return 1;
        int n = 1;
        if (needs_carries()) n += 1 + (cross_needed() ? 1 : 0) + (local_p2 ? 0 : (d_needs() ? 1 : 0) + (x_needs() ? 1 : 0));
        return n * nslices;
    }

    // difference basis of the fused carry algebra (fdiff_fwd / fdiff_inv in fused.cuh)
    static void diff_fwd(HT* c) { for (int m = 1; m < R; ++m) for (int k = R - 1; k >= m; --k) c[k] = c[k - 1] - c[k]; }
    static void diff_inv(HT* c) { for (int m = R - 1; m >= 1; --m) for (int k = m; k < R; ++k) c[k] = c[k - 1] - c[k]; }
    // M <- D M D^-1 for every R x R block of `mats`;  G <- G D^-1 for every row of `rows` ([n][R])
    static void conjugate_blocks(std::vector<HT>& mats)
    {
        for (size_t b = 0; b + (size_t)R * R <= mats.size(); b += (size_t)R * R) {
            HT out[R * R];
            for (int j = 0; j < R; ++j) {
                HT v[R], w[R];
                for (int k = 0; k < R; ++k) v[k] = (HT)(k == j ? 1 : 0);
                diff_inv(v);
                for (int i = 0; i < R; ++i) {
                    HT acc = (HT)0;
                    for (int k = 0; k < R; ++k) acc = acc + mats[b + i * R + k] * v[k];
                    w[i] = acc;
                }
                diff_fwd(w);
                for (int i = 0; i < R; ++i) out[i * R + j] = w[i];
            }
            std::copy(out, out + R * R, mats.begin() + b);
        }
    }
    static void right_multiply_rows(std::vector<HT>& rows)
    {
        HT dinv[R * R];                                   // column j = D^-1 e_j
        for (int j = 0; j < R; ++j) {
            HT v[R];
            for (int k = 0; k < R; ++k) v[k] = (HT)(k == j ? 1 : 0);
            diff_inv(v);
            for (int k = 0; k < R; ++k) dinv[k * R + j] = v[k];
        }
        for (size_t b = 0; b + (size_t)R <= rows.size(); b += (size_t)R) {
            HT out[R];
            for (int j = 0; j < R; ++j) {
                HT acc = (HT)0;
                for (int k = 0; k < R; ++k) acc = acc + rows[b + k] * dinv[k * R + j];
                out[j] = acc;
            }
            std::copy(out, out + R, rows.begin() + b);
        }
    }

    // product of the per-tile transition matrices over each segment of FCHAIN_L tiles, segments in scan order
    static std::vector<HT> build_fseg(const DimTables<HT>& tb, const std::vector<HostScan>& scans, int nb, int nseg, int L)
    {
        const int S = (int)scans.size();
        std::vector<HT> out((size_t)S * nseg * R * R, (HT)0), tmp;
        for (int s = 0; s < S; ++s)
            for (int gs = 0; gs < nseg; ++gs) {
                const int g = scans[s].causal ? gs : nseg - 1 - gs;          // memory segment
                const int j0 = g * L, j1 = std::min(nb, j0 + L);
                std::vector<HT> acc((size_t)R * R, (HT)0);
                for (int k = 0; k < R; ++k) acc[k * R + k] = (HT)1;
                for (int t = 0; t < j1 - j0; ++t) {
                    const int j = scans[s].causal ? j0 + t : j1 - 1 - t;
                    const int var = nb == 1 ? V_SINGLE : (j == 0 ? V_FIRST : (j == nb - 1 ? V_LAST : V_INTERIOR));
                    matmul_rr<HT>(tmp, &tb.P[((size_t)var * S + s) * R * R], acc.data(), R);
                    acc = tmp;
                }
                std::copy(acc.begin(), acc.end(), out.begin() + ((size_t)s * nseg + gs) * R * R);
            }
        return out;
    }

    int init(int64_t Nx, int64_t Nd, int64_t No, bool clamp_)
    {
        clamp = clamp_;
        std::memset(&fp, 0, sizeof(fp));
        fp.Nx = Nx; fp.Nd = Nd; fp.No = No;
        fp.nbx = gx.nb; fp.nbd = gd.nb;
        fp.clamp = clamp ? 1 : 0;
        fp.x_lo_closed = gx.lo_closed; fp.x_hi_closed = gx.hi_closed;
        fp.d_lo_closed = gd.lo_closed; fp.d_hi_closed = gd.hi_closed;
        fp.mx = gx.nscans; fp.md = gd.nscans;
        fp.nlx = Nd * No; fp.nly = Nx * No;
        double gain = 1.0; uint32_t gain_u = 1u;
        auto fill = [&](FusedScanTab<CT, R>& tab, const std::vector<HostScan>& sc) {
            for (size_t s = 0; s < sc.size(); ++s) {
                tab.causal[s] = sc[s].causal;
                const std::vector<HT> c = coeff_vec<HT>(sc[s], R, true);
                for (int k = 0; k <= R; ++k) tab.a[s][k] = (CT)c[k];
                gain *= (double)sc[s].coeff[0];
                gain_u *= cvt_coeff<uint32_t>(sc[s].coeff[0]);
            }
        };
        fill(fp.sx, sx); fill(fp.sd, sd);
        fp.gain = std::is_same<CT, float>::value ? (CT)gain : (CT)gain_u;
        Lx = fchain_pick_L(gx.nb, fp.mx);
        Ld = fchain_pick_L(gd.nb, fp.md);
        nsegx = (gx.nb + Lx - 1) / Lx;
        nsegd = (gd.nb + Ld - 1) / Ld;

        // short memory: every entry of every tile transition matrix (difference basis, fp64) is below 1e-10, so
        // P * carry is far below the last bit of the tail it would be added to (RFB_NO_LOCAL_CARRY=1: always chain)
        auto short_memory = [&](const DimTables<HT>& tb, int nscans, int nb) {
            if (!std::is_same<CT, float>::value || nscans < 1 || nscans > 2 || nb > 65535) return false;
            if (const char* e = getenv("RFB_NO_LOCAL_CARRY")) if (atoi(e)) return false;
            for (const HT& v : tb.P) if (!(std::fabs((double)v) < 1e-10)) return false;
            return true;
        };
        if (fp.mx > 0) {
            build_dim_tables<HT>(tx_tab, sx, gx, R, clamp, 0, 0, true, true, ts);
            conjugate_blocks(tx_tab.P); conjugate_blocks(tx_tab.M);
            local_x = short_memory(tx_tab, fp.mx, gx.nb);
            CUDA_TRY((upload<HT, TT>(dPx, tx_tab.P)));
            CUDA_TRY((upload<HT, TT>(dMx, tx_tab.M)));
            CUDA_TRY((upload<HT, TT>(dL, tx_tab.L)));
            CUDA_TRY((upload<HT, TT>(dPsegx, build_fseg(tx_tab, sx, gx.nb, nsegx, Lx))));
            const size_t n = (size_t)fp.mx * R * gx.nb * fp.nlx * sizeof(CT);
            CUDA_TRY(TX.alloc(n)); CUDA_TRY(CX.alloc(n));
            CUDA_TRY(cudaMemset(CX.p, 0, n));
        }
        if (fp.md > 0) {
            build_dim_tables<HT>(td_tab, sd, gd, R, clamp, 0, 0, true, true, ts);
            conjugate_blocks(td_tab.P); conjugate_blocks(td_tab.M); right_multiply_rows(td_tab.G);
            local_d = short_memory(td_tab, fp.md, gd.nb);
            CUDA_TRY((upload<HT, TT>(dPd, td_tab.P)));
            CUDA_TRY((upload<HT, TT>(dMd, td_tab.M)));
            CUDA_TRY((upload<HT, TT>(dG, td_tab.G)));
            CUDA_TRY((upload<HT, TT>(dPsegd, build_fseg(td_tab, sd, gd.nb, nsegd, Ld))));
            const size_t n = (size_t)fp.md * R * gd.nb * fp.nly * sizeof(CT);
            CUDA_TRY(TY.alloc(n)); CUDA_TRY(CY.alloc(n));
            CUDA_TRY(cudaMemset(CY.p, 0, n));
            if (d_open()) {
                const size_t m = (size_t)fp.md * R * fp.nly * sizeof(CT);
                CUDA_TRY(dExt.alloc(m)); CUDA_TRY(dTailOut.alloc(m));
                CUDA_TRY(cudaMemset(dExt.p, 0, m));
                resolver.reset(new ShardResolver<CT, R>());
                int rc = resolver->init(sd, gd.n, clamp, true);
                if (rc) return rc;
                std::vector<HT> Wv;
                build_strip_response<HT>(Wv, sd, gd.n, ts, gd.nb, gd.lo_closed, gd.hi_closed, R, clamp, true);
                conjugate_blocks(Wv);
                CUDA_TRY((upload<HT, TT>(dW, Wv)));
                // tiles whose response blocks all round to zero in the compute type are skipped by the kernel
                std::vector<unsigned char> act((size_t)fp.md * gd.nb, 0);
                for (int s2 = 0; s2 < fp.md; ++s2)
                    for (int q = 0; q <= s2; ++q)
                        for (int j = 0; j < gd.nb; ++j)
                            for (int i = 0; i < R * R; ++i)
                                if ((TT)Wv[((((size_t)s2 * fp.md + q) * gd.nb + j) * R * R) + i] != (TT)0) act[(size_t)s2 * gd.nb + j] = 1;
                CUDA_TRY(dWact.alloc(act.size()));
                CUDA_TRY(cudaMemcpy(dWact.p, act.data(), act.size(), cudaMemcpyHostToDevice));
            }
        }
        if (cross_needed()) {
            const size_t n = (size_t)gx.nb * gd.nb * No * fp.mx * R * sdk() * sizeof(CT);
            CUDA_TRY(dA.alloc(n));
            CUDA_TRY(cudaMemset(dA.p, 0, n));
        }
        fp.TX = (CT*)TX.p; fp.CX = (const CT*)CX.p; fp.TY = (CT*)TY.p; fp.CY = (const CT*)CY.p;
        // M[0 -> 1] per tile variant as kernel constants (short-memory pass 2)
        for (int var = 0; var < V_COUNT; ++var)
            for (int i = 0; i < R * R; ++i) {
                fp.Mlx[var][i] = fp.mx > 1 ? (CT)tx_tab.M[((((size_t)var * fp.mx + 0) * fp.mx + 1) * R * R) + i] : (CT)0;
                fp.Mld[var][i] = fp.md > 1 ? (CT)td_tab.M[((((size_t)var * fp.md + 0) * fp.md + 1) * R * R) + i] : (CT)0;
            }
        // short memory in every scanned dimension: pass 2 can derive the carries entering a tile from the tails of the
        // neighbouring tiles (FusedParams::local), so that no carry kernel is launched at all -- P1, the cross residual
        // A (from the tails), P2.  Needs full tiles, an unsharded pass, and the larger staging area must still allow
        // three CTAs per SM.  OFF by default (RFB_LOCAL_P2=1 turns it on): measured on 8192^2, the carry stage shrinks
        // from 30 to 19 us per image but pass 2 -- which is bound by the bytes its three CTAs per SM keep in flight --
        // pays for every instruction in front of its scans: 94.6 -> 112 us for one image (RFB_LOCAL_P2=1: pass 2 also applies
        // the cross residual G_row * A), or 101 us with a 31 us cross kernel that corrects the x tails in place
        // (RFB_LOCAL_P2=2); 156.5 / 158.3 against 157 us per image in a stack of eight.  Checked against the chained carries
        // (tests/test_fused_gpu.py).
        {
            const bool off = !(getenv("RFB_LOCAL_P2") && atoi(getenv("RFB_LOCAL_P2")) != 0);
            local_mode = (getenv("RFB_LOCAL_P2") && atoi(getenv("RFB_LOCAL_P2")) == 2) ? 2 : 1;
            const size_t smem = fused_tile_smem_bytes(ts, fused_p2_carry_words(fp.mx, fp.md, R, ts, 1, sdk()));
            const int per_sm = ts == 128 ? 3 : 6;
            local_p2 = !off && needs_carries() && !d_open() && (fp.mx == 0 || local_x || gx.nb == 1) && (fp.md == 0 || local_d || gd.nb == 1) &&
                       fp.mx <= 2 && fp.md <= 2 && Nx % ts == 0 && Nd % ts == 0 && (smem + 1024) * per_sm <= 233472;
            // the same conditions allow both sweeps in ONE launch (fused_stream_kernel: pass 2 follows pass 1 a few tile
            // rows behind and reads its input from L2: 8 B/sample of HBM traffic instead of 12).  OFF by default
            // (RFB_STREAM=1 turns it on): measured, an SM that holds a mix of pass-1 and pass-2 items loses more to
            // instruction-cache misses than the filter gains in traffic -- 8192^2: 250 us against 156 per image; stacks
            // that fit L2 tie (4096^2: 79.2 against 78.9 us), see fused.cuh.  Checked against the two sweeps in
            // tests/test_fused_gpu.py.
            const bool want = getenv("RFB_STREAM") && atoi(getenv("RFB_STREAM")) != 0;
            const int64_t rows = No * gd.nb;
            stream_ok = want && std::is_same<CT, float>::value && ts == 128 && cross_needed() && !d_open() && local_x && local_d &&
                        fp.mx <= 2 && fp.md <= 2 && Nx % ts == 0 && Nd % ts == 0 && (smem + 1024) * 3 <= 233472 &&
                        rows < (1 << 24) && (rows + 64) * (2 * (int64_t)gx.nb + (gx.nb + 3) / 4) < 0x7fffffffLL;
            if (stream_ok) {
                CUDA_TRY(dSync.alloc((size_t)(2 + 2 * rows) * sizeof(unsigned)));
                CUDA_TRY(dStreamErr.alloc(sizeof(unsigned)));
                CUDA_TRY(cudaMemset(dStreamErr.p, 0, sizeof(unsigned)));
                // the cross residuals of a row wait for pass 1 of two rows further on, pass 2 for the residuals: keep
                // about a wave of resident CTAs (3 per SM) between producer and consumer so that waiting is rare,
                // and no more (every tile row pass 2 lags behind has to stay in L2, beside what pass 2 writes)
                int sms = 148, dev = 0;
                if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const int step = 2 * gx.nb + (gx.nb + 3) / 4;
                const int gap = (3 * sms + step - 1) / step;
                lag_a = 2 + gap; lag_p = lag_a + gap;
                if (const char* e = getenv("RFB_STREAM_LAG_A")) lag_a = std::max(2, atoi(e));
                if (const char* e = getenv("RFB_STREAM_LAG_P")) lag_p = atoi(e);
                lag_p = std::max(lag_p, lag_a);
            }
        }
        return init_pipeline();
    }

    // ---- stack slices: images [oa, ob) of the stack; oa == ob == 0 means the whole stack ----
    int64_t sl_a = 0, sl_b = 0;
    void set_slice(int64_t oa, int64_t ob) { sl_a = oa; sl_b = ob; fp.o0 = oa; fp.No_launch = ob - oa; }

    int run_tails(const void* in, void* out, cudaStream_t st) override
    {
        if (!needs_carries()) return RF_OK;
        cudaEvent_t ev = timer ? timer->begin(st, ST_TAILS) : nullptr;
        fp.reverse = 0;
        CUDA_TRY((FLaunch<CT, R>::tile(fp, in, out, FMODE_P1, ts, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    /*
     * Pipelined stack (optional, RFB_PIPE_SLICES=n; OFF by default).  The carry stage is a chain of dependent,
     * latency-bound launches that use a fraction of the machine (19 % of a C3 step when it runs alone).  The stack is
     * cut into slices of whole images; the tile kernels of all slices run back to back on the caller's stream, the
     * carry stage of slice i runs on a high-priority side stream beside the tile kernels of slice i+1 (P1) / i-1 (P2):
     *     stream : P1(0) P1(1) P2(0) P1(2) P2(1) ... P2(n-1)
     *     side   :       C(0)        C(1)  ...   C(n-1)
     * Measured on B200 (scripts/c3_pipe.py, 4 x 8192^2): 158.9 us per image unsliced, 170.1 with 2 slices, 188.1 with 4
     * -- the results are bit-identical but nothing overlaps: three resident tile CTAs use the whole register file of
     * an SM (3 x 128 x 168), so a chain block (256 threads x 88 registers, 58-70 KB) only fits once the tile kernel
     * drains, and every extra kernel boundary costs a ramp.  Kept for the record and for smaller tile kernels.
     */
    int nslices = 1;
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> ev_tails, ev_carries;
    ~FusedPass()
    {
        for (cudaEvent_t e : ev_tails) cudaEventDestroy(e);
        for (cudaEvent_t e : ev_carries) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
    }
    int init_pipeline()
    {
        nslices = 1;
        if (fp.No < 2 || !needs_carries() || d_open()) return RF_OK;
        const int64_t tiles_per_image = (int64_t)gx.nb * gd.nb;
        int want = 1;
        if (const char* e = getenv("RFB_PIPE_SLICES")) want = atoi(e);
        if (want < 2) return RF_OK;
        // slice boundaries must fall on chain blocks (32 lines) and cross-residual blocks (4 tiles)
        if ((fp.Nx % 32) || (fp.Nd % 32) || (tiles_per_image % 4)) return RF_OK;
        nslices = (int)std::min<int64_t>(want, fp.No);
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));        // hi = greatest priority (numerically lowest)
        CUDA_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
        ev_tails.resize(nslices); ev_carries.resize(nslices);
        for (int i = 0; i < nslices; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&ev_tails[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_carries[i], cudaEventDisableTiming));
        }
        return RF_OK;
    }
    const uint32_t* error_flag() const override { return stream_ok ? (const uint32_t*)dStreamErr.p : nullptr; }

    // both sweeps in one launch
    int run_stream(const void* in, void* out, cudaStream_t st)
    {
        if constexpr (std::is_same<CT, float>::value && R <= 4) {
            FStreamParams<CT, R> sp;
            std::memset(&sp, 0, sizeof(sp));
            sp.t = fp;
            sp.t.reverse = 0; sp.t.local = 1; sp.t.prefetch = 0; sp.t.o0 = 0; sp.t.No_launch = 0;
            sp.t.Mx = (const TT*)dMx.p; sp.t.Md = (const TT*)dMd.p; sp.t.sdk = sdk();
            sp.t.A = (const CT*)dA.p; sp.t.G = (const TT*)dG.p;
            FCrossParams<CT, R>& cr = sp.c;
            cr.CY = (const CT*)CY.p; cr.A = (CT*)dA.p; cr.L = (const TT*)dL.p;
            cr.Nx = fp.Nx; cr.Nd = fp.Nd; cr.No = fp.No; cr.nbx = gx.nb; cr.nbd = gd.nb; cr.Sx = fp.mx; cr.Sd = fp.md;
            cr.sdk = sdk(); cr.nly = fp.nly; cr.nlx = fp.nlx;
            cr.local = 1; cr.TY = (const CT*)TY.p; cr.Md = (const TT*)dMd.p; cr.TXw = (CT*)TX.p; cr.G = (const TT*)dG.p;
            for (int s2 = 0; s2 < fp.md && s2 < 2; ++s2) cr.causal_d[s2] = sd[s2].causal;
            cr.w0 = 0; cr.w1 = (int64_t)gx.nb * gd.nb * fp.No;
            const int64_t rows = fp.No * gd.nb;
            unsigned* w = (unsigned*)dSync.p;
            sp.ticket = w; sp.cnt_p1 = w + 2; sp.cnt_a = w + 2 + rows; sp.err = (unsigned*)dStreamErr.p;
            sp.rows = (int)rows; sp.na = (gx.nb + 3) / 4; sp.step = 2 * gx.nb + sp.na;
            sp.lag_a = lag_a; sp.lag_p = lag_p;
            cudaEvent_t ev = timer ? timer->begin(st, ST_FINAL) : nullptr;
            CUDA_TRY(cudaMemsetAsync(dSync.p, 0, dSync.bytes, st));
            CUDA_TRY(launch_fused_stream(sp, in, out, st));
            if (timer) timer->end(st, ev);
            return RF_OK;
        } else {
            (void)in; (void)out; (void)st;
            return fail(RF_EINTERNAL, "fused_stream_kernel is a float kernel");
        }
    }

    int run_all(const void* in, void* out, cudaStream_t st) override
    {
        if (stream_ok && nslices < 2) return run_stream(in, out, st);
        if (nslices < 2) return PassBase::run_all(in, out, st);
        auto bounds = [&](int i, int64_t& a, int64_t& b) { a = fp.No * i / nslices; b = fp.No * (i + 1) / nslices; };
        int rc = RF_OK;
        int64_t a, b;
        auto tails = [&](int i) -> int {
            bounds(i, a, b); set_slice(a, b);
            int r = run_tails(in, out, st);
            if (r) return r;
            CUDA_TRY(cudaEventRecord(ev_tails[i], st));
            return RF_OK;
        };
        if ((rc = tails(0))) { set_slice(0, 0); return rc; }
        for (int i = 0; i < nslices && !rc; ++i) {
            if (i + 1 < nslices) rc = tails(i + 1);
            if (rc) break;
            bounds(i, a, b); set_slice(a, b);
            cudaError_t e = cudaStreamWaitEvent(side, ev_tails[i], 0);
            if (e == cudaSuccess) { rc = run_carries(nullptr, nullptr, side, 0); if (rc) break; }
            if (e == cudaSuccess) e = cudaEventRecord(ev_carries[i], side);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev_carries[i], 0);
            if (e != cudaSuccess) { rc = fail(RF_ECUDA, "pipelined stack: %s", cudaGetErrorString(e)); break; }
            rc = run_final(in, out, st);
        }
        set_slice(0, 0);
        return rc;
    }

    // carries of a short-memory dimension: one streaming launch, no chain (flocal_kernel)
    int run_local(bool xdim, void* tail_out_d, cudaStream_t st)
    {
        FLocalParams<CT, R> lp;
        std::memset(&lp, 0, sizeof(lp));
        const DimGeom& g = xdim ? gx : gd;
        const auto& scans = xdim ? sx : sd;
        lp.T = (const CT*)(xdim ? TX.p : TY.p);
        lp.C = (CT*)(xdim ? CX.p : CY.p);
        lp.nl = xdim ? fp.nlx : fp.nly;
        lp.nb = g.nb; lp.S = g.nscans;
        for (int s = 0; s < g.nscans; ++s) lp.causal[s] = scans[s].causal;
        lp.M = (const TT*)(xdim ? dMx.p : dMd.p);
        lp.tail_out = xdim ? nullptr : (CT*)tail_out_d;
        if (xdim && cross_needed()) {
            lp.A = (const CT*)dA.p; lp.G = (const TT*)dG.p;
            lp.Nd = fp.Nd; lp.nbd = gd.nb; lp.Sd = fp.md; lp.ts = ts; lp.sdk = sdk();
        }
        cudaEvent_t ev = timer ? timer->begin(st, ST_CHAIN) : nullptr;
        CUDA_TRY((FLaunch<CT, R>::local(lp, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    int run_chain(bool xdim, const void* ext_d, void* tail_out_d, cudaStream_t st, bool tails_only = false)
    {
        // short-memory dimension, whole stack, nothing entering the strip (a sharded stage 2 corrects the carries
        // afterwards) and the carries wanted: no chain
        // (measured on 4 x 8192^2: 6.5 us per image against 10 for the d chain; with the cross-dimension residual folded
        // into the x tails the streaming kernel loses to the x chain, which stages the A matrices in shared memory:
        // 28 us against 13.5 -- so the x dimension keeps its chain whenever there are d scans)
        if ((xdim ? (local_x && !cross_needed()) : local_d) && !ext_d && !tails_only && sl_b == sl_a) return run_local(xdim, tail_out_d, st);
        FChainParams<CT, R> cp;
        std::memset(&cp, 0, sizeof(cp));
        const DimGeom& g = xdim ? gx : gd;
        const auto& scans = xdim ? sx : sd;
        cp.T = (const CT*)(xdim ? TX.p : TY.p);
        cp.C = (CT*)(xdim ? CX.p : CY.p);
        cp.nl = xdim ? fp.nlx : fp.nly;
        cp.nb = g.nb; cp.S = g.nscans;
        cp.nseg = xdim ? nsegx : nsegd;
        cp.L = xdim ? Lx : Ld;
        for (int s = 0; s < g.nscans; ++s) cp.causal[s] = scans[s].causal;
        cp.P = (const TT*)(xdim ? dPx.p : dPd.p);
        cp.M = (const TT*)(xdim ? dMx.p : dMd.p);
        cp.Pseg = (const TT*)(xdim ? dPsegx.p : dPsegd.p);
        cp.ext = xdim ? nullptr : (const CT*)ext_d;
        cp.tail_out = xdim ? nullptr : (CT*)tail_out_d;
        cp.sJ = cp.nl; cp.sL = 1;
        if (sl_b > sl_a) {                           // a slice of the stack: its lines only
            const int64_t per_image = xdim ? fp.Nd : fp.Nx;
            cp.l0 = sl_a * per_image; cp.l1 = sl_b * per_image;
        }
        cp.no_store = tails_only ? 1 : 0;            // stage 1 of a sharded run wants the outgoing tails only
        if (xdim && cross_needed()) {
            cp.A = (const CT*)dA.p; cp.G = (const TT*)dG.p;
            cp.Nd = fp.Nd; cp.nbd = gd.nb; cp.Sd = fp.md; cp.ts = ts; cp.sdk = sdk();
        }
        cudaEvent_t ev = timer ? timer->begin(st, ST_CHAIN) : nullptr;
        CUDA_TRY((FLaunch<CT, R>::chain(cp, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    int run_carries(const void* ext_d, void* tail_out_d, cudaStream_t st, int stage) override
    {
        if (!needs_carries()) return RF_OK;
        local_now = local_p2 && stage == 0 && !ext_d && !tail_out_d;
        if (local_now) {
            // no carry kernels: only the cross-dimension residual A, from the d tails
            if (cross_needed()) {
                FCrossParams<CT, R> cr;
                std::memset(&cr, 0, sizeof(cr));
                cr.CY = (const CT*)CY.p; cr.A = (CT*)dA.p; cr.L = (const TT*)dL.p;
                cr.Nx = fp.Nx; cr.Nd = fp.Nd; cr.No = fp.No; cr.nbx = gx.nb; cr.nbd = gd.nb; cr.Sx = fp.mx; cr.Sd = fp.md;
                cr.sdk = sdk();
                cr.nly = fp.nly; cr.nlx = fp.nlx;
                // RFB_LOCAL_P2=1: A stored, pass 2 applies it; =2: this kernel corrects the x tails in place
                cr.local = local_mode; cr.TY = (const CT*)TY.p; cr.Md = (const TT*)dMd.p;
                cr.TXw = (CT*)TX.p; cr.G = (const TT*)dG.p;
                for (int s2 = 0; s2 < fp.md && s2 < 2; ++s2) cr.causal_d[s2] = sd[s2].causal;
                if (sl_b > sl_a) { cr.w0 = sl_a * (int64_t)gx.nb * gd.nb; cr.w1 = sl_b * (int64_t)gx.nb * gd.nb; }
                cudaEvent_t ev = timer ? timer->begin(st, ST_CROSS) : nullptr;
                CUDA_TRY((FLaunch<CT, R>::cross(cr, ts, st)));
                if (timer) timer->end(st, ev);
            }
            return RF_OK;
        }
        // strip sharding: stage 1 chains the strip with zero carries entering it, keeps those carries and emits the
        // outgoing tails; stage 2 does NOT chain again -- the carries are corrected by W * ext (carry_fix_kernel)
        static const bool rechain = getenv("RFB_SHARD_RECHAIN") && atoi(getenv("RFB_SHARD_RECHAIN")) != 0;   // old scheme, for comparison
        if (stage == 2 && !rechain && d_needs() && dW.p) {
            if (ext_d) {
                cudaEvent_t ev = timer ? timer->begin(st, ST_CHAIN) : nullptr;
                carry_fix_kernel<CT, R><<<dim3((unsigned)((fp.nly + 127) / 128), (unsigned)gd.nb), 128, 0, st>>>(
                    (CT*)CY.p, (const CT*)ext_d, (const CT*)dW.p, (const unsigned char*)dWact.p, fp.md, gd.nb, fp.nly);
                CUDA_TRY(cudaGetLastError());
                if (timer) timer->end(st, ev);
            }
        } else if (d_needs()) {
            int rc = run_chain(false, ext_d, tail_out_d, st, stage == 1 && rechain);
            if (rc) return rc;
        }
        if (stage == 1) return RF_OK;                // stage 1 of a sharded run: the exchange comes next
        if (cross_needed()) {
            FCrossParams<CT, R> cr;
            std::memset(&cr, 0, sizeof(cr));
            cr.CY = (const CT*)CY.p; cr.A = (CT*)dA.p; cr.L = (const TT*)dL.p;
            cr.Nx = fp.Nx; cr.Nd = fp.Nd; cr.No = fp.No; cr.nbx = gx.nb; cr.nbd = gd.nb; cr.Sx = fp.mx; cr.Sd = fp.md;
            cr.sdk = sdk();
            cr.nly = fp.nly; cr.nlx = fp.nlx;
            if (sl_b > sl_a) { cr.w0 = sl_a * (int64_t)gx.nb * gd.nb; cr.w1 = sl_b * (int64_t)gx.nb * gd.nb; }
            cudaEvent_t ev = timer ? timer->begin(st, ST_CROSS) : nullptr;
            CUDA_TRY((FLaunch<CT, R>::cross(cr, ts, st)));
            if (timer) timer->end(st, ev);
        }
        if (x_needs()) { int rc = run_chain(true, nullptr, nullptr, st); if (rc) return rc; }
        return RF_OK;
    }

    int run_final(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = timer ? timer->begin(st, ST_FINAL) : nullptr;
        fp.reverse = needs_carries() ? 1 : 0;
        fp.local = (local_p2 && local_now) ? 1 : 0;
        fp.Mx = (const TT*)dMx.p; fp.Md = (const TT*)dMd.p; fp.sdk = sdk();
        fp.A = (cross_needed() && local_mode == 1) ? (const CT*)dA.p : nullptr; fp.G = (const TT*)dG.p;
        CUDA_TRY((FLaunch<CT, R>::tile(fp, in, out, FMODE_P2, ts, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }

    bool set_epilogue(float a_in, float a_out) override
    {
        if (!std::is_same<CT, float>::value || R > 4 || fp.mx == 0 || d_open() || local_p2) return false;
        stream_ok = false;                       // the one-launch kernel has no epilogue instantiation
        fp.epilogue = 1; fp.epi_in = (CT)a_in; fp.epi_out = (CT)a_out;
        return true;
    }
    size_t shard_tail_elems() const override { return d_open() ? (size_t)fp.md * R * fp.nly : 0; }
    int shard_resolve(const void* gathered, int nshards, int rank, cudaStream_t st) override
    {
        if (!d_open()) return RF_OK;
        return resolver->run(fp.nly, gathered, nshards, rank, dExt.p, st);
    }
    int shard_resolve_lines(const void* gathered, int nshards, int64_t nlines, void* ext_all, cudaStream_t st) override
    {
        if (!d_open()) return RF_OK;
        return resolver->run(nlines, gathered, nshards, -1, ext_all, st);
    }
    int shard_vectors() const override { return fp.md * R; }
    bool shard_neighbors_suffice() const override { return resolver && resolver->neighbors_suffice; }

    std::string describe() const override
    {
        char b[512];
        snprintf(b, sizeof(b),
                 "  fused pass view [%lld][%lld][%lld]: %dx%d register tiles, d scans %d (%d tiles) then x scans %d "
                 "(%d tiles), order<=%d, unit feed-forward (gain applied at the store), two sweeps: 12 B/sample, launches %d%s%s\n",
                 (long long)fp.No, (long long)fp.Nd, (long long)fp.Nx, ts, ts, fp.md, fp.nbd, fp.mx, fp.nbx, R,
                 launches(), fp.epilogue ? " (pointwise epilogue a*in + b*filtered fused into the store)" :
                 (stream_ok && nslices < 2) ? " -- whole-filter calls: both sweeps in ONE launch (short memory: pass 2 follows pass 1 a few tile rows behind and reads its input from L2, 8 B/sample of HBM traffic)" : "",
                 (std::string(nslices > 1 ? " (stack pipelined in " + std::to_string(nslices) + " slices: carry stages on a side stream)" : "") +
                              (local_p2 ? std::string(" (short memory: pass 2 derives its carries from the neighbouring tiles' tails, no carry kernels)") :
                               (local_d || (local_x && !cross_needed())) ? std::string(" (short-memory carries, no chain, along") + (local_d ? " d" : "") + ((local_x && !cross_needed()) ? " x" : "") + ")" : std::string(""))).c_str());
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
// long 1-D signals (apps/audio: one causal order-8 scan over 64 channels x 2^24 samples).
// The signal is viewed as rows of 128 samples; the fused tile kernels run with one thread per row
// (signal mode: a row continues the row before it), so the "tiles" of a signal are its rows and the
// chain runs along the row index.  A signal has M = N/128 rows -- far more than one chain launch
// handles -- so the chain is a hierarchy: level l chains groups of nb_l consecutive tiles and emits
// one aggregate tail per group (up-sweep), the top level sees one line per signal, and the levels
// are re-run top-down with the carry entering each group (down-sweep).  All levels are the same
// fchain_kernel with a strided layout (tile index contiguous); the transition matrix of level l+1
// is P_l ^ nb_l.
// ---------------------------------------------------------------------------------------------
template <typename CT, int R>
struct SignalPass : PassBase {
    using HT = typename std::conditional<std::is_same<CT, float>::value, double, uint32_t>::type;
    FusedParams<CT, R> fp;
    static constexpr int ts = 128;
    HostScan scan;
    int64_t nsig = 1, M = 1;                       // signals, rows per signal
    struct Level {
        int nb = 1, L = 8, nseg = 1;
        int64_t nl = 1;                            // lines = groups of nb tiles
        DevBuf T, C;                               // [R][nl * nb]   (level 0: the x tails / carries of the tile kernels)
        DevBuf dP, dM, dPseg;
    };
    std::vector<std::unique_ptr<Level>> lv;

    bool needs_carries() const override { return true; }
    bool d_open() const override { return false; }
    const void* ext_buffer() const override { return nullptr; }
    size_t shard_tail_elems() const override { return 0; }
    int shard_resolve(const void*, int, int, cudaStream_t) override { return RF_OK; }
    int shard_resolve_lines(const void*, int, int64_t, void*, cudaStream_t) override { return RF_OK; }
    int shard_vectors() const override { return 0; }
    size_t workspace() const override
    {
        size_t n = 0;
        for (auto& l : lv) n += l->T.bytes + l->C.bytes + l->dP.bytes + l->dM.bytes + l->dPseg.bytes;
        return n;
    }
    int launches() const override { return 2 + 2 * (int)lv.size() - 1; }

    // fan-out of one level: the largest divisor of m that one chain launch handles, multiples of 8 preferred
    static int pick_fanout(int64_t m)
    {
        if (m <= 64) return (int)m;
        for (int d = 64; d >= 8; d -= 8) if (m % d == 0) return d;
        for (int d = 63; d >= 2; --d) if (m % d == 0) return d;
        return 0;
    }
    static bool eligible(int64_t Nx, int64_t rows)
    {
        if (Nx % ts || Nx / ts < 2 || ((Nx / ts) * rows) % ts) return false;
        if ((Nx / ts) * rows > 0x7fffffffLL / (4 * R)) return false;        // 32-bit carry offsets
        int64_t m = Nx / ts;
        int levels = 0;
        while (m > 1) { const int f = pick_fanout(m); if (f < 2) return false; m /= f; if (++levels > 6) return false; }
        return true;
    }

    int init(int64_t Nx, int64_t rows, bool clamp)
    {
        M = Nx / ts; nsig = rows;
        std::memset(&fp, 0, sizeof(fp));
        fp.Nx = ts; fp.Nd = M * nsig; fp.No = 1;
        fp.nbx = 1; fp.nbd = (int)(fp.Nd / ts);
        fp.clamp = clamp ? 1 : 0;
        fp.x_lo_closed = fp.x_hi_closed = fp.d_lo_closed = fp.d_hi_closed = 1;
        fp.mx = 1; fp.md = 0;
        fp.nlx = fp.Nd; fp.nly = ts;
        fp.signal = 1; fp.sig_rows = M;
        fp.sx.causal[0] = scan.causal;
        const std::vector<HT> c = coeff_vec<HT>(scan, R, true);
        for (int k = 0; k <= R; ++k) fp.sx.a[0][k] = (CT)c[k];
        fp.gain = std::is_same<CT, float>::value ? (CT)(double)scan.coeff[0] : (CT)cvt_coeff<uint32_t>(scan.coeff[0]);

        // transition of one row (interior tile, unit feed-forward form), difference basis
        DimGeom g; g.n = Nx; g.t = ts; g.nb = (int)M; g.len_last = ts; g.nscans = 1; g.lo_closed = 1; g.hi_closed = 1;
        DimTables<HT> tab;
        build_dim_tables<HT>(tab, std::vector<HostScan>(1, scan), g, R, clamp, 0, 0, false, true, ts);
        std::vector<HT> P(tab.P.begin() + (size_t)V_INTERIOR * R * R, tab.P.begin() + (size_t)(V_INTERIOR + 1) * R * R);
        FusedPass<CT, R>::conjugate_blocks(P);

        int64_t m = M;
        while (true) {
            std::unique_ptr<Level> l(new Level());
            l->nb = m > 1 ? pick_fanout(m) : 1;
            l->nl = nsig * (m / l->nb);
            l->L = l->nb <= 32 ? 4 : FCHAIN_L;
            l->nseg = (l->nb + l->L - 1) / l->L;
            const size_t bytes = (size_t)R * l->nl * l->nb * sizeof(CT);
            CUDA_TRY(l->T.alloc(bytes)); CUDA_TRY(l->C.alloc(bytes));
            CUDA_TRY(cudaMemset(l->C.p, 0, bytes));
            // tables: every tile is an interior tile
            std::vector<HT> Pv, Mv((size_t)V_COUNT * R * R, (HT)0), Pseg, acc, tmp;
            for (int v = 0; v < V_COUNT; ++v) Pv.insert(Pv.end(), P.begin(), P.end());
            for (int gs = 0; gs < l->nseg; ++gs) {
                const int gmem = scan.causal ? gs : l->nseg - 1 - gs;
                const int cnt = std::min(l->nb, (gmem + 1) * l->L) - gmem * l->L;
                acc.assign((size_t)R * R, (HT)0);
                for (int k = 0; k < R; ++k) acc[k * R + k] = (HT)1;
                for (int t = 0; t < cnt; ++t) { matmul_rr<HT>(tmp, P.data(), acc.data(), R); acc = tmp; }
                Pseg.insert(Pseg.end(), acc.begin(), acc.end());
            }
            CUDA_TRY((upload<HT, CT>(l->dP, Pv)));
            CUDA_TRY((upload<HT, CT>(l->dM, Mv)));
            CUDA_TRY((upload<HT, CT>(l->dPseg, Pseg)));
            // next level: one tile = one group of this level
            acc.assign((size_t)R * R, (HT)0);
            for (int k = 0; k < R; ++k) acc[k * R + k] = (HT)1;
            for (int t = 0; t < l->nb; ++t) { matmul_rr<HT>(tmp, P.data(), acc.data(), R); acc = tmp; }
            P = acc;
            m /= l->nb;
            lv.push_back(std::move(l));
            if (m <= 1) break;
        }
        fp.TX = (CT*)lv[0]->T.p; fp.CX = (const CT*)lv[0]->C.p;
        return RF_OK;
    }

    int run_tails(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = timer ? timer->begin(st, ST_TAILS) : nullptr;
        fp.reverse = 0;
        CUDA_TRY((FLaunch<CT, R>::tile(fp, in, out, FMODE_P1, ts, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }
    int run_level(int i, const void* ext, void* tail_out, bool store, cudaStream_t st)
    {
        Level& l = *lv[i];
        FChainParams<CT, R> cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.T = (const CT*)l.T.p; cp.C = (CT*)l.C.p;
        cp.nl = l.nl; cp.nb = l.nb; cp.S = 1; cp.nseg = l.nseg; cp.L = l.L;
        cp.causal[0] = scan.causal;
        cp.P = (const CT*)l.dP.p; cp.M = (const CT*)l.dM.p; cp.Pseg = (const CT*)l.dPseg.p;
        cp.ext = (const CT*)ext; cp.tail_out = (CT*)tail_out;
        cp.sJ = 1; cp.sL = l.nb; cp.uniform = 1; cp.no_store = store ? 0 : 1;
        cudaEvent_t ev = timer ? timer->begin(st, ST_CHAIN) : nullptr;
        CUDA_TRY((FLaunch<CT, R>::chain(cp, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }
    int run_carries(const void*, void*, cudaStream_t st, int) override
    {
        const int top = (int)lv.size() - 1;
        int rc;
        for (int i = 0; i < top; ++i)                         // up-sweep: group aggregates
            if ((rc = run_level(i, nullptr, lv[i + 1]->T.p, false, st))) return rc;
        if ((rc = run_level(top, nullptr, nullptr, true, st))) return rc;
        for (int i = top - 1; i >= 0; --i)                    // down-sweep: carries entering each group
            if ((rc = run_level(i, lv[i + 1]->C.p, nullptr, true, st))) return rc;
        return RF_OK;
    }
    int run_final(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = timer ? timer->begin(st, ST_FINAL) : nullptr;
        fp.reverse = 0;
        CUDA_TRY((FLaunch<CT, R>::tile(fp, in, out, FMODE_P2, ts, st)));
        if (timer) timer->end(st, ev);
        return RF_OK;
    }
    std::string describe() const override
    {
        std::string fan;
        for (auto& l : lv) fan += (fan.empty() ? "" : "x") + std::to_string(l->nb);
        char b[512];
        snprintf(b, sizeof(b),
                 "  signal pass: %lld signals x %lld rows of %d samples (thread per row, register tiles), 1 %s scan of "
                 "order<=%d, carry hierarchy %s, launches %d\n",
                 (long long)nsig, (long long)M, ts, scan.causal ? "causal" : "anticausal", R, fan.c_str(), launches());
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
// single-pass passes (kernels in lookback.cuh): every tile is read once and written once, the
// inter-tile carries travel by decoupled look-back inside the one kernel.  Eligible: at most one
// scan per dimension (any direction), full tiles, unsharded.
// ---------------------------------------------------------------------------------------------
template <typename CT, int R>
struct LookbackBase : PassBase {
    using HT = typename std::conditional<std::is_same<CT, float>::value, double, uint32_t>::type;
    DevBuf dCtl;                                   // words [0,1] 64-bit ticket counter (never reset), [2] error flag
    bool needs_carries() const override { return false; }
    bool d_open() const override { return false; }
    const void* ext_buffer() const override { return nullptr; }
    size_t shard_tail_elems() const override { return 0; }
    int shard_resolve(const void*, int, int, cudaStream_t) override { return RF_OK; }
    int shard_resolve_lines(const void*, int, int64_t, void*, cudaStream_t) override { return RF_OK; }
    int shard_vectors() const override { return 0; }
    int launches() const override { return 1; }
    int run_tails(const void*, void*, cudaStream_t) override { return RF_OK; }
    int run_carries(const void*, void*, cudaStream_t, int) override { return RF_OK; }
    const uint32_t* error_flag() const override { return dCtl.p ? (const uint32_t*)dCtl.p + 2 : nullptr; }
    int init_ctl()
    {
        CUDA_TRY(dCtl.alloc(4 * sizeof(uint32_t)));
        CUDA_TRY(cudaMemset(dCtl.p, 0, 4 * sizeof(uint32_t)));
        return RF_OK;
    }
    // transition of one tile of `ts` samples (unit feed-forward form), raw basis
    static std::vector<HT> tile_transition(const HostScan& sc, int ts, bool clamp)
    {
        DimGeom g; g.n = 4 * (int64_t)ts; g.t = ts; g.nb = 4; g.len_last = ts; g.nscans = 1; g.lo_closed = 1; g.hi_closed = 1;
        DimTables<HT> tab;
        build_dim_tables<HT>(tab, std::vector<HostScan>(1, sc), g, R, clamp, 0, 0, false, true, ts);
        return std::vector<HT>(tab.P.begin() + (size_t)V_INTERIOR * R * R, tab.P.begin() + (size_t)(V_INTERIOR + 1) * R * R);
    }
    static std::vector<HT> identity()
    {
        std::vector<HT> m((size_t)R * R, (HT)0);
        for (int k = 0; k < R; ++k) m[k * R + k] = (HT)1;
        return m;
    }
    static std::vector<HT> mat_pow(const std::vector<HT>& m, int64_t e)       // m^e by squaring
    {
        std::vector<HT> acc = identity(), base = m, tmp;
        while (e > 0) {
            if (e & 1) { matmul_rr<HT>(tmp, base.data(), acc.data(), R); acc = tmp; }
            e >>= 1;
            if (e) { matmul_rr<HT>(tmp, base.data(), base.data(), R); base = tmp; }
        }
        return acc;
    }
    // m^0 .. m^(n-1), concatenated
    static std::vector<HT> mat_powers(const std::vector<HT>& m, int n)
    {
        std::vector<HT> out, acc = identity(), tmp;
        for (int j = 0; j < n; ++j) {
            out.insert(out.end(), acc.begin(), acc.end());
            matmul_rr<HT>(tmp, m.data(), acc.data(), R); acc = tmp;
        }
        return out;
    }
    template <size_t N>
    static void to_device_consts(CT (&dst)[N], std::vector<HT> m)           // difference basis, rounded once
    {
        FusedPass<CT, R>::conjugate_blocks(m);
        for (size_t i = 0; i < N && i < m.size(); ++i) dst[i] = (CT)m[i];
    }
};

template <typename CT, int R>
struct LookbackPass : LookbackBase<CT, R> {
    using Base = LookbackBase<CT, R>;
    using HT = typename Base::HT;
    LBTileParams<CT, R> lp;
    int ts = 128;
    std::vector<HostScan> sx, sd;
    DevBuf dPpow[2], dRec[2];

    size_t workspace() const override
    {
        size_t n = this->dCtl.bytes;
        for (int i = 0; i < 2; ++i) n += dPpow[i].bytes + dRec[i].bytes;
        return n;
    }
    int init_dim(LBDim<CT, R>& dm, const std::vector<HostScan>& sc, int nb, int64_t ntiles, bool clamp, int slot, double& gain, uint32_t& gain_u)
    {
        std::memset(&dm, 0, sizeof(dm));
        dm.nscan = (int)sc.size();
        if (sc.empty()) return RF_OK;
        dm.causal = sc[0].causal;
        const std::vector<HT> c = coeff_vec<HT>(sc[0], R, true);
        for (int k = 0; k <= R; ++k) dm.a[k] = (CT)c[k];
        gain *= (double)sc[0].coeff[0];
        gain_u *= cvt_coeff<uint32_t>(sc[0].coeff[0]);
        const std::vector<HT> P = Base::tile_transition(sc[0], ts, clamp);
        Base::to_device_consts(dm.P, P);
        std::vector<HT> pw = Base::mat_powers(P, std::max(nb, 1));
        FusedPass<CT, R>::conjugate_blocks(pw);
        CUDA_TRY((upload<HT, CT>(dPpow[slot], pw)));
        // per tile and line: one vector of (R + 2) / 3 self-validating 16-byte chunks (epoch 0 = nothing)
        const size_t n = (size_t)ntiles * ts * ((R + 2) / 3) * 16;
        CUDA_TRY(dRec[slot].alloc(n));
        CUDA_TRY(cudaMemset(dRec[slot].p, 0, n));
        dm.Ppow = (const CT*)dPpow[slot].p; dm.rec = dRec[slot].p;
        return RF_OK;
    }
    int init(int64_t Nx, int64_t Nd, int64_t No, bool clamp)
    {
        std::memset(&lp, 0, sizeof(lp));
        lp.Nx = Nx; lp.Nd = Nd; lp.No = No;
        lp.nbx = (int)(Nx / ts); lp.nbd = (int)(Nd / ts);
        lp.clamp = clamp ? 1 : 0;
        int rc = this->init_ctl();
        if (rc) return rc;
        const int64_t ntiles = (int64_t)lp.nbx * lp.nbd * No;
        double gain = 1.0; uint32_t gain_u = 1u;
        if ((rc = init_dim(lp.x, sx, lp.nbx, ntiles, clamp, 0, gain, gain_u))) return rc;
        if ((rc = init_dim(lp.d, sd, lp.nbd, ntiles, clamp, 1, gain, gain_u))) return rc;
        lp.gain = std::is_same<CT, float>::value ? (CT)gain : (CT)gain_u;
        lp.ticket = (unsigned long long*)this->dCtl.p; lp.err = (uint32_t*)this->dCtl.p + 2;
        // tiles of an image are handed out along anti-diagonals of the scan-order grid (the tiles a tile waits for are
        // then a whole diagonal older; measured on the 8192^2 table: 134 vs 143 us); RFB_LB_ORDER=rows: row-major
        lp.rows_first = (getenv("RFB_LB_ORDER") && !strcmp(getenv("RFB_LB_ORDER"), "rows")) ? 1 : 0;
        // optional L2 prefetch distance in tickets (RFB_LB_PREFETCH; measured: no gain, off)
        lp.prefetch = 0;
        if (const char* e = getenv("RFB_LB_PREFETCH")) lp.prefetch = atoi(e);
        return RF_OK;
    }
    int run_final(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = this->timer ? this->timer->begin(st, ST_FINAL) : nullptr;
        CUDA_TRY((LBLaunch<CT, R>::tile(lp, in, out, ts, st)));
        if (this->timer) this->timer->end(st, ev);
        return RF_OK;
    }
    std::string describe() const override
    {
        char b[512];
        snprintf(b, sizeof(b),
                 "  single-pass look-back view [%lld][%lld][%lld]: %dx%d register tiles, d scans %d (%d tiles, %s) then x scans %d "
                 "(%d tiles, %s), order<=%d, unit feed-forward, 1 launch, 8 B/sample\n",
                 (long long)lp.No, (long long)lp.Nd, (long long)lp.Nx, ts, ts, lp.d.nscan, lp.nbd, lp.d.causal ? "causal" : "anticausal",
                 lp.x.nscan, lp.nbx, lp.x.causal ? "causal" : "anticausal", R);
        return b;
    }
};

template <typename CT, int R>
struct SignalLookbackPass : LookbackBase<CT, R> {
    using Base = LookbackBase<CT, R>;
    using HT = typename Base::HT;
    static constexpr int ts = 128;
    LBSignalParams<CT, R> sp;
    HostScan scan;
    int64_t nsig = 1, M = 1;
    DevBuf dPlane, dQpow, dRec;

    int tile_rows = 128;                                                      // rows of 128 samples per CTA
    // rows per CTA: the largest of 128 / 64 / 32 that cuts a signal into whole tiles (measured on C4: 2.07 / 2.31 /
    // 2.83 ms -- the look-back cost per tile outweighs the extra CTAs per SM; RFB_LB_ROWS overrides)
    static int pick_tile_rows(int64_t Nx)
    {
        if (const char* e = getenv("RFB_LB_ROWS")) { const int v = atoi(e); if ((v == 128 || v == 64 || v == 32) && Nx % ((int64_t)v * ts) == 0) return v; }
        for (int v : { 128, 64, 32 }) if (Nx % ((int64_t)v * ts) == 0) return v;
        return 0;
    }
    static bool eligible(int64_t Nx, int64_t rows)
    {
        const int tr = pick_tile_rows(Nx);
        if (!tr) return false;                                                // whole tiles of tr rows x 128 samples per signal
        const int64_t tiles = (Nx / ((int64_t)tr * ts)) * rows;
        return tiles > 0 && tiles <= 0x7fffffffLL && (Nx / ts) * rows <= 0x7fffffffLL;
    }
    size_t workspace() const override { return this->dCtl.bytes + dPlane.bytes + dQpow.bytes + dRec.bytes; }

    // [n][R][R] matrices -> [R*R][32] (lane fastest), difference basis
    static std::vector<HT> lane_table(std::vector<HT> mats)
    {
        FusedPass<CT, R>::conjugate_blocks(mats);
        std::vector<HT> out((size_t)R * R * 32, (HT)0);
        for (int l = 0; l < 32; ++l)
            for (int i = 0; i < R * R; ++i) out[(size_t)i * 32 + l] = mats[(size_t)l * R * R + i];
        return out;
    }
    int init(int64_t Nx, int64_t rows, bool clamp)
    {
        M = Nx / ts; nsig = rows;
        std::memset(&sp, 0, sizeof(sp));
        tile_rows = pick_tile_rows(Nx);
        sp.rows = M * nsig;
        sp.tile_rows = tile_rows;
        sp.tiles_per_signal = (int)(M / tile_rows);
        sp.causal = scan.causal; sp.clamp = clamp ? 1 : 0;
        const std::vector<HT> c = coeff_vec<HT>(scan, R, true);
        for (int k = 0; k <= R; ++k) sp.a[k] = (CT)c[k];
        sp.gain = std::is_same<CT, float>::value ? (CT)(double)scan.coeff[0] : (CT)cvt_coeff<uint32_t>(scan.coeff[0]);
        int rc = this->init_ctl();
        if (rc) return rc;
        const std::vector<HT> P = Base::tile_transition(scan, ts, clamp);     // one row of 128 samples
        for (int i = 0; i < 5; ++i) Base::to_device_consts(sp.Pstep[i], Base::mat_pow(P, 1 << i));
        Base::to_device_consts(sp.Pwarp, Base::mat_pow(P, 32));
        const std::vector<HT> Q = Base::mat_pow(P, tile_rows);
        Base::to_device_consts(sp.Q, Q);
        Base::to_device_consts(sp.Q32, Base::mat_pow(Q, 32));
        CUDA_TRY((upload<HT, CT>(dPlane, lane_table(Base::mat_powers(P, 32)))));
        CUDA_TRY((upload<HT, CT>(dQpow, lane_table(Base::mat_powers(Q, 32)))));
        const int64_t ntiles = sp.rows / tile_rows;
        CUDA_TRY(dRec.alloc((size_t)ntiles * LB_SIGNAL_REC_CHUNKS * 16));
        CUDA_TRY(cudaMemset(dRec.p, 0, (size_t)ntiles * LB_SIGNAL_REC_CHUNKS * 16));
        sp.Plane = (const CT*)dPlane.p; sp.Qpow = (const CT*)dQpow.p;
        sp.rec = dRec.p;
        sp.ticket = (unsigned long long*)this->dCtl.p; sp.err = (uint32_t*)this->dCtl.p + 2;
        // ---- short-memory specialisations, decided from fp64 bounds (RFB_NO_SHORT_MEMORY=1: off) ----
        sp.pass0_first_chunk = 0; sp.depth1 = 0;
        if (std::is_same<CT, float>::value && !(getenv("RFB_NO_SHORT_MEMORY") && atoi(getenv("RFB_NO_SHORT_MEMORY")))) {
            // (a) weight of sample i of a row (scan order) in the row's tail: L[k][i]; chunks whose weights are all below
            //     1e-12 of the largest one are skipped by pass 0
            std::vector<HT> cf = coeff_vec<HT>(scan, R, true);
            double lmax = 0.0, cmax[4] = { 0, 0, 0, 0 };
            for (int i = 0; i < ts; ++i) {
                std::vector<HT> v(ts, (HT)0), hh(R, (HT)0);
                v[i] = (HT)1;
                sim_scan<HT>(v, ts, hh, cf, true, false, true);
                for (int k = 0; k < R; ++k) { const double a = std::fabs((double)hh[k]); lmax = std::max(lmax, a); cmax[i / 32] = std::max(cmax[i / 32], a); }
            }
            while (sp.pass0_first_chunk < 3 && cmax[sp.pass0_first_chunk] < 1e-12 * lmax) ++sp.pass0_first_chunk;
            // (b) transition of a whole tile
            double qmax = 0.0;
            for (const HT& q : Q) qmax = std::max(qmax, std::fabs((double)q));
            std::vector<HT> Qd = Q;
            FusedPass<CT, R>::conjugate_blocks(Qd);
            for (const HT& q : Qd) qmax = std::max(qmax, std::fabs((double)q));
            if (qmax < 1e-12) sp.depth1 = 1;
        }
        sp.prefetch = 0;                                                    // optional L2 prefetch distance (measured: no gain)
        if (const char* e = getenv("RFB_LB_PREFETCH")) sp.prefetch = atoi(e);
        return RF_OK;
    }
    int run_final(const void* in, void* out, cudaStream_t st) override
    {
        cudaEvent_t ev = this->timer ? this->timer->begin(st, ST_FINAL) : nullptr;
        CUDA_TRY((LBLaunch<CT, R>::signal(sp, in, out, st)));
        if (this->timer) this->timer->end(st, ev);
        return RF_OK;
    }
    std::string describe() const override
    {
        char b[512];
        snprintf(b, sizeof(b),
                 "  single-pass look-back signal pass: %lld signals x %lld rows of %d samples (thread per row, %d rows per CTA, "
                 "%d CTAs per signal), 1 %s scan of order<=%d, 1 launch, 8 B/sample%s%s\n",
                 (long long)nsig, (long long)M, ts, tile_rows, sp.tiles_per_signal, scan.causal ? "causal" : "anticausal", R,
                 sp.pass0_first_chunk ? (", short memory: the tail pass skips " + std::to_string(32 * sp.pass0_first_chunk) + " samples per row").c_str() : "",
                 sp.depth1 ? ", carries from the previous tile only" : "");
        return b;
    }
};

// narrow integer types are widened to the 32-bit compute ring on entry and truncated on exit
// (arithmetic mod 2^16 / 2^8 is a quotient of arithmetic mod 2^32, so this is exact)
template <typename ST>
__global__ void widen_kernel(const ST* __restrict__ in, uint32_t* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint32_t)in[i];
}
template <typename ST>
__global__ void narrow_kernel(const uint32_t* __restrict__ in, ST* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (ST)in[i];
}


// ---------------------------------------------------------------------------------------------
// pointwise linear stencil of one array (rf_stencil_execute): box-filter finite differencing etc.
// ---------------------------------------------------------------------------------------------
struct StencilParams {
    int ndim, ntaps;
    float post_scale;
    int64_t extent[RF_MAX_DIMS];
    int64_t total;
    rf_tap tap[RF_MAX_TAPS];
};

template <typename CT>
__global__ void __launch_bounds__(256) stencil_kernel(const __grid_constant__ StencilParams p, const CT* __restrict__ in,
                                                      const CT* __restrict__ in2, CT* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c[RF_MAX_DIMS], r = i;
#pragma unroll
        for (int d = 0; d < RF_MAX_DIMS; ++d) {
            if (d < p.ndim) { c[d] = r % p.extent[d]; r /= p.extent[d]; } else c[d] = 0;
        }
        CT acc = (CT)0;
        for (int t = 0; t < p.ntaps; ++t) {
            int64_t idx = 0, stride = 1;
#pragma unroll
            for (int d = 0; d < RF_MAX_DIMS; ++d) {
                if (d < p.ndim) {
                    int64_t v = c[d] + p.tap[t].offset[d];
                    v = min(v, (int64_t)p.tap[t].hi[d]);              // clamp(a, lo, hi) = max(min(a, hi), lo)
                    v = max(v, (int64_t)p.tap[t].lo[d]);
                    v = max((int64_t)0, min(v, p.extent[d] - 1));
                    idx += v * stride; stride *= p.extent[d];
                }
            }
            const CT w = std::is_same<CT, float>::value ? (CT)p.tap[t].weight : (CT)(int32_t)lrintf(p.tap[t].weight);
            acc = acc + w * __ldg((p.tap[t].source ? in2 : in) + idx);
        }
        out[i] = std::is_same<CT, float>::value ? acc * (CT)p.post_scale : acc * (CT)(int32_t)lrintf(p.post_scale);
    }
}

// 1-D / 2-D arrays: 32-bit coordinates, four consecutive outputs per thread (one 128-bit store when the row
// length allows), rows on blockIdx.y -- neighbouring threads read neighbouring words of every tap, so the taps
// are coalesced and mostly served by L1/L2
template <typename CT>
__global__ void __launch_bounds__(256) stencil2d_kernel(const __grid_constant__ StencilParams p, const CT* __restrict__ in,
                                                        const CT* __restrict__ in2, CT* __restrict__ out)
{
    // a block owns 1024 consecutive x of a row; thread t the samples xb + t + 256 e.  Every load / store instruction of a
    // warp therefore covers 32 consecutive words whatever the tap offset is (a tap at x + B reads unaligned addresses:
    // with 4 consecutive samples per thread each instruction touched four lines and the kernel was bound by L1
    // wavefronts -- 52 us for the 4-tap box stencil on 4096^2, against 20 us of HBM time)
    const int W = (int)p.extent[0], Hh = p.ndim > 1 ? (int)p.extent[1] : 1;
    const int xb = blockIdx.x * 1024 + threadIdx.x;
    if (xb >= W) return;
    const CT scale = std::is_same<CT, float>::value ? (CT)p.post_scale : (CT)(int32_t)lrintf(p.post_scale);
    // a block walks a band of consecutive rows: a tap at y + B finds the row again 2B+1 rows later, in L2
    const int band = (Hh + (int)gridDim.y - 1) / (int)gridDim.y;
    const int y_end = min(Hh, ((int)blockIdx.y + 1) * band);
    for (int y = (int)blockIdx.y * band; y < y_end; ++y) {
        CT acc[4] = { (CT)0, (CT)0, (CT)0, (CT)0 };
        for (int t = 0; t < p.ntaps; ++t) {
            const rf_tap& tp = p.tap[t];
            int yy = max(min(y + tp.offset[1], tp.hi[1]), tp.lo[1]);
            yy = max(0, min(yy, Hh - 1));
            const CT* row = (tp.source ? in2 : in) + (size_t)yy * W;
            const CT w = std::is_same<CT, float>::value ? (CT)tp.weight : (CT)(int32_t)lrintf(tp.weight);
            const int xs = xb + tp.offset[0];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int xx = max(min(xs + 256 * e, tp.hi[0]), tp.lo[0]);
                xx = max(0, min(xx, W - 1));
                acc[e] = acc[e] + w * __ldg(row + xx);
            }
        }
        CT* o = out + (size_t)y * W + xb;
#pragma unroll
        for (int e = 0; e < 4; ++e) if (xb + 256 * e < W) o[256 * e] = acc[e] * scale;
    }
}

} // namespace rfb

// ---------------------------------------------------------------------------------------------
// the opaque plan
// ---------------------------------------------------------------------------------------------
using namespace rfb;

struct rf_plan {
    rf_desc desc;
    int device = 0;          // the device that was current at rf_plan_create: workspace, tables and launches belong to it
    int R = 1;
    bool is_float = true;
    size_t elem_bytes = 4;
    int64_t total = 1;
    std::vector<std::unique_ptr<PassBase>> passes;
    int shard_pass = -1;
    std::string text;
    DevBuf stage;            // 32-bit staging for 8/16-bit integer filters
    StageTimer timer;
    // host-buffer path (rf_plan_execute_host / _batch): three device image buffers cycled through
    // upload -> filter (in place) -> download on three streams, created on first use
    struct HostPipe {
        static constexpr int NB = 3;
        DevBuf buf[NB];
        cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
        cudaEvent_t up[NB] = {}, done[NB] = {}, down[NB] = {};
        bool ready = false;
        ~HostPipe()
        {
            if (!ready) return;
            for (int i = 0; i < NB; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(done[i]); cudaEventDestroy(down[i]); }
            cudaStreamDestroy(s_up); cudaStreamDestroy(s_run); cudaStreamDestroy(s_down);
        }
    } pipe;
};

static int widen_in(rf_plan* plan, const void* in_dev, cudaStream_t st)
{
    const int64_t n = plan->total;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (plan->elem_bytes == 2) widen_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)in_dev, (uint32_t*)plan->stage.p, n);
    else                       widen_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)in_dev, (uint32_t*)plan->stage.p, n);
    CUDA_TRY(cudaGetLastError());
    return RF_OK;
}
static int narrow_out(rf_plan* plan, void* out_dev, cudaStream_t st)
{
    const int64_t n = plan->total;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    if (plan->elem_bytes == 2) narrow_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint32_t*)plan->stage.p, (uint16_t*)out_dev, n);
    else                       narrow_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint32_t*)plan->stage.p, (uint8_t*)out_dev, n);
    CUDA_TRY(cudaGetLastError());
    return RF_OK;
}

static int round_order(int r)
{
    const int opts[] = { 1, 2, 3, 4, 8, 16, 32 };
    for (int o : opts) if (r <= o) return o;
    return -1;
}

template <typename CT, int R>
static int make_pass(rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                     int64_t Nx, int64_t Nd, int64_t No, int tx, int td, bool fused, bool d_is_shard)
{
    const rf_desc& d = plan->desc;
    auto ps = std::unique_ptr<Pass<CT, R>>(new (std::nothrow) Pass<CT, R>());
    if (!ps) return fail(RF_ENOMEM, "out of host memory");
    ps->dtype = d.dtype;
    ps->sx = sx; ps->sd = sd; ps->fused = fused;
    auto geom = [](DimGeom& g, int64_t n, int t, int nscans) {
        g.n = n; g.t = t; g.nb = (int)((n + t - 1) / t); g.len_last = (int)(n - (int64_t)(g.nb - 1) * t);
        g.nscans = nscans; g.lo_closed = 1; g.hi_closed = 1;
    };
    geom(ps->gx, Nx, tx, (int)sx.size());
    geom(ps->gd, Nd, td, (int)sd.size());
    if (d_is_shard) { ps->gd.lo_closed = d.opt.open_lo ? 0 : 1; ps->gd.hi_closed = d.opt.open_hi ? 0 : 1; }
    // signal mode: too few rows to give every thread of a CTA its own line
    const bool signal = sd.empty() && !sx.empty() && (Nd * No) < TILE && ps->gx.nb > 1;
    int rc = ps->init(Nx, Nd, No, signal, d.border == RF_BORDER_CLAMP);
    if (rc) return rc;
    plan->passes.push_back(std::move(ps));
    return RF_OK;
}

template <typename CT, int R>
static int make_fused_pass(rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                           int64_t Nx, int64_t Nd, int64_t No, int ts, bool d_is_shard)
{
    const rf_desc& d = plan->desc;
    auto ps = std::unique_ptr<FusedPass<CT, R>>(new (std::nothrow) FusedPass<CT, R>());
    if (!ps) return fail(RF_ENOMEM, "out of host memory");
    ps->sx = sx; ps->sd = sd; ps->ts = ts;
    auto geom = [ts](DimGeom& g, int64_t n, int nscans) {
        g.n = n; g.t = ts; g.nb = (int)((n + ts - 1) / ts); g.len_last = (int)(n - (int64_t)(g.nb - 1) * ts);   // ragged: partial last tile
        g.nscans = nscans; g.lo_closed = 1; g.hi_closed = 1;
    };
    geom(ps->gx, Nx, (int)sx.size());
    geom(ps->gd, Nd, (int)sd.size());
    if (d_is_shard) { ps->gd.lo_closed = d.opt.open_lo ? 0 : 1; ps->gd.hi_closed = d.opt.open_hi ? 0 : 1; }
    int rc = ps->init(Nx, Nd, No, d.border == RF_BORDER_CLAMP);
    if (rc) return rc;
    plan->passes.push_back(std::move(ps));
    return RF_OK;
}

template <typename CT, int R>
static int make_signal_pass(rf_plan* plan, const HostScan& sc, int64_t Nx, int64_t rows)
{
    auto ps = std::unique_ptr<SignalPass<CT, R>>(new (std::nothrow) SignalPass<CT, R>());
    if (!ps) return fail(RF_ENOMEM, "out of host memory");
    ps->scan = sc;
    int rc = ps->init(Nx, rows, plan->desc.border == RF_BORDER_CLAMP);
    if (rc) return rc;
    plan->passes.push_back(std::move(ps));
    return RF_OK;
}

template <typename CT, int R>
static int make_lookback_pass(rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                              int64_t Nx, int64_t Nd, int64_t No, int ts)
{
    auto ps = std::unique_ptr<LookbackPass<CT, R>>(new (std::nothrow) LookbackPass<CT, R>());
    if (!ps) return fail(RF_ENOMEM, "out of host memory");
    ps->sx = sx; ps->sd = sd; ps->ts = ts;
    int rc = ps->init(Nx, Nd, No, plan->desc.border == RF_BORDER_CLAMP);
    if (rc) return rc;
    plan->passes.push_back(std::move(ps));
    return RF_OK;
}
template <typename CT>
static int make_lookback_pass_R(rf_plan* plan, int R, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                                int64_t Nx, int64_t Nd, int64_t No, int ts)
{
    switch (R) {
    case 1: return make_lookback_pass<CT, 1>(plan, sx, sd, Nx, Nd, No, ts);
    case 2: return make_lookback_pass<CT, 2>(plan, sx, sd, Nx, Nd, No, ts);
    case 3: return make_lookback_pass<CT, 3>(plan, sx, sd, Nx, Nd, No, ts);
    case 4: return make_lookback_pass<CT, 4>(plan, sx, sd, Nx, Nd, No, ts);
    }
    return fail(RF_EUNSUPPORTED, "the single-pass tile kernel supports orders <= 4");
}
template <typename CT, int R>
static int make_signal_lookback_pass(rf_plan* plan, const HostScan& sc, int64_t Nx, int64_t rows)
{
    auto ps = std::unique_ptr<SignalLookbackPass<CT, R>>(new (std::nothrow) SignalLookbackPass<CT, R>());
    if (!ps) return fail(RF_ENOMEM, "out of host memory");
    ps->scan = sc;
    int rc = ps->init(Nx, rows, plan->desc.border == RF_BORDER_CLAMP);
    if (rc) return rc;
    plan->passes.push_back(std::move(ps));
    return RF_OK;
}
template <typename CT>
static int make_signal_lookback_pass_R(rf_plan* plan, int R, const HostScan& sc, int64_t Nx, int64_t rows)
{
    switch (R) {
    case 1: return make_signal_lookback_pass<CT, 1>(plan, sc, Nx, rows);
    case 2: return make_signal_lookback_pass<CT, 2>(plan, sc, Nx, rows);
    case 3: return make_signal_lookback_pass<CT, 3>(plan, sc, Nx, rows);
    case 4: return make_signal_lookback_pass<CT, 4>(plan, sc, Nx, rows);
    case 8: return make_signal_lookback_pass<CT, 8>(plan, sc, Nx, rows);
    }
    return fail(RF_EUNSUPPORTED, "the single-pass signal kernel supports orders <= 8");
}

// single-pass kernels allowed?  (RF_ENGINE_TWOPASS / RFB_NO_LOOKBACK=1 keep the multi-pass kernels, for comparison)
static bool lookback_allowed(const rf_plan* plan)
{
    const rf_options& opt = plan->desc.opt;
    if (opt.engine == RF_ENGINE_GENERIC || opt.engine == RF_ENGINE_TWOPASS || opt.honor_tile) return false;
    if (opt.open_lo || opt.open_hi) return false;
    if (opt.epilogue) return false;               // the fused epilogue lives in the two-sweep kernels' pass 2
    if (const char* e = getenv("RFB_NO_LOOKBACK")) if (atoi(e)) return false;
    return true;
}
static bool unit_ff_ok(const rf_plan* plan, const HostScan& h)
{
    if (plan->is_float) {
        const double inv = 1.0 / (double)h.coeff[0];
        return h.coeff[0] != 0.f && std::isfinite((float)inv);
    }
    return cvt_coeff<uint32_t>(h.coeff[0]) == 1u;
}
// tile size of the single-pass 2-D kernel for this pass, or 0: at most one scan per dimension (8 B/sample instead
// of the 12 B/sample of the two-sweep scheme -- fewer HBM bytes, so it is preferred whenever it applies)
static int lookback_tile_size(const rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                              int64_t Nx, int64_t Nd, int64_t No)
{
    if (!lookback_allowed(plan) || plan->R > 4) return 0;
    if (sx.size() > 1 || sd.size() > 1 || (sx.empty() && sd.empty())) return 0;
    double gain = 1.0;
    for (const auto* v : { &sx, &sd })
        for (const HostScan& h : *v) { if (!unit_ff_ok(plan, h)) return 0; gain *= std::fabs((double)h.coeff[0]); }
    if (plan->is_float && !(gain > 1e-24 && gain < 1e24)) return 0;      // unit feed-forward form: see fused_tile_size
    const char* force = getenv("RFB_LB_TS");
    for (int ts : { 128, 64 }) {
        if (force && atoi(force) != ts) continue;
        if (Nx % ts || Nd % ts) continue;
        const int64_t ntiles = (Nx / ts) * (Nd / ts) * No;
        if (ntiles > 0x7fffffffLL || Nx / ts > 0x7fffLL || Nd / ts > 0x7fffLL) continue;
        // less than one generation of resident CTAs (3 of 128x128 per SM): 64x64 tiles fill the machine better
        // (measured: 2048^2 14.4 vs 15.0 us, 4096^2 45.5 vs 44.1 us, 8192^2 142 vs 130 us)
        if (!force && ts == 128 && ntiles < 3 * 148 && Nx % 64 == 0 && Nd % 64 == 0) continue;
        return ts;
    }
    return 0;
}
static bool signal_lookback_eligible(const rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                                     int64_t Nx, int64_t rows)
{
    if (!lookback_allowed(plan) || plan->R > 8) return false;
    if (!sd.empty() || sx.size() != 1 || !unit_ff_ok(plan, sx[0])) return false;
    return SignalLookbackPass<float, 1>::eligible(Nx, rows);
}

// a pass with a single scan along the contiguous dimension of long lines: the signal pass (orders <= 8)
static bool signal_eligible(const rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                            int64_t Nx, int64_t rows)
{
    const rf_options& opt = plan->desc.opt;
    if (opt.engine == RF_ENGINE_GENERIC || opt.honor_tile || opt.epilogue) return false;
    if (!sd.empty() || sx.size() != 1 || plan->R > 8) return false;
    if (const char* e = getenv("RFB_NO_SIGNAL")) if (atoi(e)) return false;
    const HostScan& h = sx[0];
    if (plan->is_float) {
        const double inv = 1.0 / (double)h.coeff[0];
        if (h.coeff[0] == 0.f || !std::isfinite((float)inv)) return false;
    } else if (cvt_coeff<uint32_t>(h.coeff[0]) != 1u) return false;
    switch (plan->R) {
    case 1: return SignalPass<float, 1>::eligible(Nx, rows);
    case 2: return SignalPass<float, 2>::eligible(Nx, rows);
    case 3: return SignalPass<float, 3>::eligible(Nx, rows);
    case 4: return SignalPass<float, 4>::eligible(Nx, rows);
    case 8: return SignalPass<float, 8>::eligible(Nx, rows);
    }
    return false;
}

template <typename CT>
static int make_signal_pass_R(rf_plan* plan, int R, const HostScan& sc, int64_t Nx, int64_t rows)
{
    switch (R) {
    case 1: return make_signal_pass<CT, 1>(plan, sc, Nx, rows);
    case 2: return make_signal_pass<CT, 2>(plan, sc, Nx, rows);
    case 3: return make_signal_pass<CT, 3>(plan, sc, Nx, rows);
    case 4: return make_signal_pass<CT, 4>(plan, sc, Nx, rows);
    case 8: return make_signal_pass<CT, 8>(plan, sc, Nx, rows);
    }
    return fail(RF_EUNSUPPORTED, "the signal pass supports orders <= 8");
}

// tile size of the fused fast path for this pass, or 0 when the pass must take the generic engine
static int fused_tile_size(const rf_plan* plan, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                           int64_t Nx, int64_t Nd, int64_t No)
{
    const rf_options& opt = plan->desc.opt;
    if (opt.engine == RF_ENGINE_GENERIC || opt.honor_tile) return 0;
    if (plan->R > 4) return 0;
    if (sx.size() > (size_t)FMAX_SCANS || sd.size() > (size_t)FMAX_SCANS || (sx.empty() && sd.empty())) return 0;
    for (const auto* v : { &sx, &sd })
        for (const HostScan& h : *v) {
            if (plan->is_float) {
                const double inv = 1.0 / (double)h.coeff[0];
                if (h.coeff[0] == 0.f || !std::isfinite((float)inv)) return 0;
            } else if (cvt_coeff<uint32_t>(h.coeff[0]) != 1u) return 0;
        }
    // the fused kernels run every scan with unit feed-forward and apply the product of the feed-forward coefficients
    // once at the store: the intermediates grow by prod |1 / b0| (two fused wide Gaussians reach 1e36).  Outside a safe
    // range the pass takes the generic engine, which keeps every scan in its own scale.
    if (plan->is_float) {
        double gain = 1.0;
        for (const auto* v : { &sx, &sd })
            for (const HostScan& h : *v) gain *= std::fabs((double)h.coeff[0]);
        if (!(gain > 1e-24 && gain < 1e24)) return 0;
    }
    const char* force = getenv("RFB_FUSED_TS");          // development knob: force a tile size
    static const bool ragged_ok = !(getenv("RFB_NO_RAGGED") && atoi(getenv("RFB_NO_RAGGED")) != 0);
    const bool sharded = opt.open_lo || opt.open_hi;
    for (int ts : { 128, 64 }) {
        if (force && atoi(force) != ts) continue;
        if (Nx % ts || Nd % ts) {
            // ragged extents: the last tile of a dimension is partial (TMA zero-fills / clips it, the kernels
            // mask the padding).  Needs a 16-byte row pitch for the tensor map, one image per launch along d
            // (a partial tile of a stack would reach into the next image) and unsharded strips.
            if (!ragged_ok || sharded || Nx % 4 || Nx < ts || Nd < ts) continue;
            if (Nd % ts && No != 1) continue;
        }
        const int64_t nbx = (Nx + ts - 1) / ts, nbd = (Nd + ts - 1) / ts;
        if (!sx.empty() && nbx > 16 * FCHAIN_L) continue;
        if (!sd.empty() && nbd > 16 * FCHAIN_L) continue;
        if (nbx * nbd * No > 0x7fffffffLL) continue;
        if (!fchain_fits(plan->R, (int)sx.size(), (int)sd.size(), (int)nbx, (int)nbd)) continue;
        if ((int64_t)std::max(sx.size(), sd.size()) * plan->R * std::max(nbx * Nd, nbd * Nx) * No > 0x7fffffffLL) continue;   // 32-bit carry offsets
        // few tiles per SM: 64x64 tiles balance the machine better; with first-order scans (summed-area tables) their
        // extra tails are cheap, so the switch comes later (measured: C2 4096^2 SAT 45.2 -> 41.8 us)
        if (!force && ts == 128 && nbx * nbd * No < (plan->R == 1 ? 8 : 2) * 148 && Nx >= 64 && Nd >= 64 &&
            (Nx % 64 == 0 || Nx % 128 != 0) && (Nd % 64 == 0 || Nd % 128 != 0) &&
            (sx.empty() || (Nx + 63) / 64 <= 16 * FCHAIN_L) && (sd.empty() || (Nd + 63) / 64 <= 16 * FCHAIN_L))
            continue;                       // small problem: smaller tiles fill the machine better
        return ts;
    }
    return 0;
}

template <typename CT>
static int make_fused_pass_R(rf_plan* plan, int R, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                             int64_t Nx, int64_t Nd, int64_t No, int ts, bool d_is_shard)
{
    switch (R) {
    case 1: return make_fused_pass<CT, 1>(plan, sx, sd, Nx, Nd, No, ts, d_is_shard);
    case 2: return make_fused_pass<CT, 2>(plan, sx, sd, Nx, Nd, No, ts, d_is_shard);
    case 3: return make_fused_pass<CT, 3>(plan, sx, sd, Nx, Nd, No, ts, d_is_shard);
    case 4: return make_fused_pass<CT, 4>(plan, sx, sd, Nx, Nd, No, ts, d_is_shard);
    }
    return fail(RF_EUNSUPPORTED, "the fused path supports orders <= 4");
}

template <typename CT>
static int make_pass_R(rf_plan* plan, int R, const std::vector<HostScan>& sx, const std::vector<HostScan>& sd,
                       int64_t Nx, int64_t Nd, int64_t No, int tx, int td, bool fused, bool d_is_shard)
{
    switch (R) {
    case 1:  return make_pass<CT, 1>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 2:  return make_pass<CT, 2>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 3:  return make_pass<CT, 3>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 4:  return make_pass<CT, 4>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 8:  return make_pass<CT, 8>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 16: return make_pass<CT, 16>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    case 32: return make_pass<CT, 32>(plan, sx, sd, Nx, Nd, No, tx, td, fused, d_is_shard);
    }
    return fail(RF_EUNSUPPORTED, "unsupported padded order %d", R);
}

static size_t dtype_bytes(int dt)
{
    switch (dt) {
    case RF_F32: case RF_I32: case RF_U32: return 4;
    case RF_I16: case RF_U16: return 2;
    case RF_I8:  case RF_U8:  return 1;
    }
    return 0;
}

#pragma GCC visibility push(default)
extern "C" {

const char* rf_version(void) { return "recfilter_b200 0.1 (sm_100a)"; }
const char* rf_last_error(void) { return g_err; }

int rf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rf_set_device(int device)
{
    CUDA_TRY(cudaSetDevice(device));
    return RF_OK;
}

int rf_plan_create(const rf_desc* desc, rf_plan** out)
{
    if (!desc || !out) return fail(RF_EINVAL, "null argument");
    *out = nullptr;
    if (desc->ndim < 1 || desc->ndim > RF_MAX_DIMS) return fail(RF_EINVAL, "ndim must be 1..%d", RF_MAX_DIMS);
    if (desc->nscans < 0 || desc->nscans > RF_MAX_SCANS) return fail(RF_EINVAL, "nscans must be 0..%d", RF_MAX_SCANS);
    const size_t eb = dtype_bytes(desc->dtype);
    if (!eb) return fail(RF_EINVAL, "unknown dtype %d", desc->dtype);
    if (desc->border != RF_BORDER_ZERO && desc->border != RF_BORDER_CLAMP) return fail(RF_EINVAL, "unknown border %d", desc->border);
    int64_t total = 1;
    for (int d = 0; d < desc->ndim; ++d) {
        if (desc->extent[d] < 0) return fail(RF_EINVAL, "negative extent in dimension %d", d);
        total *= desc->extent[d];
    }
    int maxr = 1;
    for (int s = 0; s < desc->nscans; ++s) {
        const rf_scan& sc = desc->scans[s];
        // lib/recfilter.cpp:296-300 (unknown dimension), :274-278 (needs feed forward + feedback)
        if (sc.dim < 0 || sc.dim >= desc->ndim) return fail(RF_EINVAL, "scan %d: dimension %d is not a filter dimension", s, sc.dim);
        if (sc.order < 1 || sc.order > RF_MAX_ORDER) return fail(RF_EINVAL, "scan %d: order must be 1..%d", s, RF_MAX_ORDER);
        maxr = std::max(maxr, sc.order);
    }
    if (rf_device_count() <= 0) return fail(RF_ENODEVICE, "no CUDA device available: this engine has no CPU fallback");

    std::unique_ptr<rf_plan> plan(new (std::nothrow) rf_plan());
    if (!plan) return fail(RF_ENOMEM, "out of host memory");
    plan->desc = *desc;
    CUDA_TRY(cudaGetDevice(&plan->device));
    plan->R = round_order(maxr);
    plan->is_float = desc->dtype == RF_F32;
    plan->elem_bytes = eb;
    plan->total = total;

    const rf_options& opt = desc->opt;
    auto tile_of = [&](int d) {
        int t = TILE;
        if (opt.honor_tile && opt.tile[d] > 0) t = std::min(opt.tile[d], TILE);
        return t;
    };

    if (total > 0 && desc->nscans > 0) {
        // group scans by dimension, keeping the add_filter order inside a dimension
        // (scans of different dimensions commute: lib/split.cpp:207-242)
        std::vector<std::vector<HostScan>> by_dim(desc->ndim);
        for (int s = 0; s < desc->nscans; ++s) {
            HostScan h; h.causal = desc->scans[s].causal ? 1 : 0; h.order = desc->scans[s].order;
            std::memcpy(h.coeff, desc->scans[s].coeff, sizeof(h.coeff));
            by_dim[desc->scans[s].dim].push_back(h);
        }
        for (int d = 0; d < desc->ndim; ++d)
            if ((int)by_dim[d].size() > MAX_SCANS_DIM)
                return fail(RF_EUNSUPPORTED, "more than %d scans along dimension %d", MAX_SCANS_DIM, d);
        const int shard_dim = (opt.open_lo || opt.open_hi) ? opt.shard_dim : -1;
        if (shard_dim == 0) return fail(RF_EUNSUPPORTED, "sharding along the contiguous dimension is not supported");
        if (shard_dim >= desc->ndim) return fail(RF_EINVAL, "shard_dim out of range");

        const std::vector<HostScan> none;
        const int R = plan->R;
        bool fuse01 = desc->ndim >= 2 && !by_dim[0].empty() && !by_dim[1].empty() && opt.fuse_dims != 0;
        if (fuse01 && (size_t)by_dim[0].size() * by_dim[1].size() * R * R > 2048) fuse01 = false;

        auto add = [&](const std::vector<HostScan>& sx, const std::vector<HostScan>& sd, int64_t Nx, int64_t Nd,
                       int64_t No, int tx, int td, bool fused, bool shard) -> int {
            // fewest HBM bytes first: the single-pass look-back kernels move 8 B/sample, the two-sweep kernels 12
            if (!shard) {
                if (const int lts = lookback_tile_size(plan.get(), sx, sd, Nx, Nd, No))
                    return plan->is_float ? make_lookback_pass_R<float>(plan.get(), R, sx, sd, Nx, Nd, No, lts)
                                          : make_lookback_pass_R<uint32_t>(plan.get(), R, sx, sd, Nx, Nd, No, lts);
            }
            const int fts = fused_tile_size(plan.get(), sx, sd, Nx, Nd, No);
            // several scans along long contiguous lines that no tile engine fuses (apps/audio/audio_filter_biquads.cpp:
            // up to a dozen causal biquads on one signal): one single-pass signal kernel per scan, 8 B/sample each, instead
            // of the generic engine (12 B/sample and a launch sequence per scan: 2.2 ms for a 4096-sample signal)
            if (!fts && !shard && sd.empty() && sx.size() > 1) {
                bool each = true;
                for (const HostScan& one : sx)
                    if (!signal_lookback_eligible(plan.get(), std::vector<HostScan>(1, one), sd, Nx, Nd * No)) each = false;
                if (each) {
                    for (const HostScan& one : sx) {
                        const int rc = plan->is_float ? make_signal_lookback_pass_R<float>(plan.get(), R, one, Nx, Nd * No)
                                                      : make_signal_lookback_pass_R<uint32_t>(plan.get(), R, one, Nx, Nd * No);
                        if (rc) return rc;
                    }
                    return RF_OK;
                }
            }
            if (!fts && !shard && signal_lookback_eligible(plan.get(), sx, sd, Nx, Nd * No))
                return plan->is_float ? make_signal_lookback_pass_R<float>(plan.get(), R, sx[0], Nx, Nd * No)
                                      : make_signal_lookback_pass_R<uint32_t>(plan.get(), R, sx[0], Nx, Nd * No);
            if (!fts && !shard && signal_eligible(plan.get(), sx, sd, Nx, Nd * No)) {
                return plan->is_float ? make_signal_pass_R<float>(plan.get(), R, sx[0], Nx, Nd * No)
                                      : make_signal_pass_R<uint32_t>(plan.get(), R, sx[0], Nx, Nd * No);
            }
            if (!fts && (opt.engine == RF_ENGINE_FUSED || opt.engine == RF_ENGINE_TWOPASS))
                return fail(RF_EUNSUPPORTED, "engine=fused requested but a pass is not eligible (needs order <= 4, a row pitch that is a multiple of 16 bytes, "
                            "extents of at least one tile, non-zero float / unit integer feed-forward)");
            int rc;
            if (fts) rc = plan->is_float ? make_fused_pass_R<float>(plan.get(), R, sx, sd, Nx, Nd, No, fts, shard)
                                         : make_fused_pass_R<uint32_t>(plan.get(), R, sx, sd, Nx, Nd, No, fts, shard);
            else     rc = plan->is_float ? make_pass_R<float>(plan.get(), R, sx, sd, Nx, Nd, No, tx, td, fused, shard)
                                         : make_pass_R<uint32_t>(plan.get(), R, sx, sd, Nx, Nd, No, tx, td, fused, shard);
            if (rc == RF_OK && shard) plan->shard_pass = (int)plan->passes.size() - 1;
            return rc;
        };

        int first_d = 0;
        if (fuse01) {
            int64_t No = 1;
            for (int d = 2; d < desc->ndim; ++d) No *= desc->extent[d];
            int rc = add(by_dim[0], by_dim[1], desc->extent[0], desc->extent[1], No, tile_of(0), tile_of(1), true,
                         shard_dim == 1);
            if (rc) return rc;
            first_d = 2;
        } else if (!by_dim[0].empty()) {
            int64_t rows = 1;
            for (int d = 1; d < desc->ndim; ++d) rows *= desc->extent[d];
            int rc = add(by_dim[0], none, desc->extent[0], rows, 1, tile_of(0), TILE, false, false);
            if (rc) return rc;
            first_d = 1;
        } else {
            first_d = 1;
        }
        for (int d = first_d; d < desc->ndim; ++d) {
            if (by_dim[d].empty()) continue;
            int64_t inner = 1, outer = 1;
            for (int e = 0; e < d; ++e) inner *= desc->extent[e];
            for (int e = d + 1; e < desc->ndim; ++e) outer *= desc->extent[e];
            int rc = add(none, by_dim[d], inner, desc->extent[d], outer, TILE, tile_of(d), false, shard_dim == d);
            if (rc) return rc;
        }
        if (shard_dim >= 0 && plan->shard_pass < 0)
            return fail(RF_EINVAL, "shard_dim %d has no scans: shard it as independent batches instead", shard_dim);
    }

    if (opt.epilogue && total > 0) {
        // apps/usm/unsharp_mask_optimized.cpp:61-66 merged into the filter: only where the whole filter is ONE two-sweep pass
        if (plan->passes.size() != 1 || eb != 4 || plan->shard_pass >= 0 || !plan->passes[0]->set_epilogue(opt.epi_in, opt.epi_out))
            return fail(RF_EUNSUPPORTED, "the fused epilogue needs a float filter that is one fused two-sweep pass with scans along dimension 0");
    }
    if (eb < 4 && total > 0 && !plan->passes.empty()) {
        if (plan->shard_pass >= 0) return fail(RF_EUNSUPPORTED, "sharded plans need a 32-bit element type");
        CUDA_TRY(plan->stage.alloc((size_t)total * 4));
    }

    char head[256];
    snprintf(head, sizeof(head), "recfilter_b200 plan: %d-D, %lld samples, dtype %d, border %s, %d scans, %d passes\n",
             desc->ndim, (long long)total, desc->dtype, desc->border ? "clamp" : "zero", desc->nscans,
             (int)plan->passes.size());
    plan->text = head;
    for (auto& p : plan->passes) { plan->text += p->describe(); p->timer = &plan->timer; }
    *out = plan.release();
    return RF_OK;
}

void rf_plan_destroy(rf_plan* plan) { delete plan; }

size_t rf_plan_workspace_bytes(const rf_plan* plan)
{
    size_t n = 0;
    if (plan) { for (auto& p : plan->passes) n += p->workspace(); n += plan->stage.bytes; }
    return n;
}

int rf_plan_num_launches(const rf_plan* plan)
{
    int n = 0;
    if (plan) { for (auto& p : plan->passes) n += p->launches(); if (plan->stage.p) n += 2; }
    return n;
}

int rf_plan_describe(const rf_plan* plan, char* buf, size_t n)
{
    if (!plan || !buf || n == 0) return fail(RF_EINVAL, "null argument");
    snprintf(buf, n, "%s", plan->text.c_str());
    return RF_OK;
}

// a plan may only run on the device it was created on
static int check_device(const rf_plan* plan)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != plan->device)
        return fail(RF_EINVAL, "the plan was created on device %d but device %d is current (rf_set_device before the call)", plan->device, dev);
    return RF_OK;
}

static int copy_through(rf_plan* plan, const void* in_dev, void* out_dev, cudaStream_t st)
{
    if (in_dev != out_dev && plan->total > 0)
        CUDA_TRY(cudaMemcpyAsync(out_dev, in_dev, (size_t)plan->total * plan->elem_bytes, cudaMemcpyDeviceToDevice, st));
    return RF_OK;
}

int rf_plan_execute(rf_plan* plan, const void* in_dev, void* out_dev, void* stream)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (plan->total == 0) return RF_OK;
    if (!in_dev || !out_dev) return fail(RF_EINVAL, "null buffer");
    if (int drc = check_device(plan)) return drc;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->passes.empty()) return copy_through(plan, in_dev, out_dev, st);
    const void* src = in_dev;
    void* dst = out_dev;
    int rc;
    if (plan->stage.p) {
        if ((rc = widen_in(plan, in_dev, st))) return rc;
        src = plan->stage.p; dst = plan->stage.p;
    }
    for (auto& p : plan->passes) {
        if ((rc = p->run_all(src, dst, st))) return rc;
        src = dst;
    }
    if (plan->stage.p) return narrow_out(plan, out_dev, st);
    return RF_OK;
}

static int host_pipe_init(rf_plan* plan, int nbuf)
{
    rf_plan::HostPipe& hp = plan->pipe;
    const size_t bytes = (size_t)plan->total * plan->elem_bytes;
    if (!hp.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_up, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_down, cudaStreamNonBlocking));
        for (int i = 0; i < rf_plan::HostPipe::NB; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&hp.up[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hp.done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hp.down[i], cudaEventDisableTiming));
        }
        hp.ready = true;
    }
    for (int i = 0; i < nbuf && i < rf_plan::HostPipe::NB; ++i)
        if (!hp.buf[i].p) CUDA_TRY(hp.buf[i].alloc(bytes));
    return RF_OK;
}

int rf_plan_execute_host_batch(rf_plan* plan, int n, const void* const* in_host, void* const* out_host)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (n < 0 || (n > 0 && (!in_host || !out_host))) return fail(RF_EINVAL, "bad argument");
    if (plan->total == 0 || n == 0) return RF_OK;
    for (int i = 0; i < n; ++i)
        if (!in_host[i] || !out_host[i]) return fail(RF_EINVAL, "null buffer");
    constexpr int NB = rf_plan::HostPipe::NB;
    int rc = host_pipe_init(plan, n < NB ? n : NB);
    if (rc) return rc;
    rf_plan::HostPipe& hp = plan->pipe;
    const size_t bytes = (size_t)plan->total * plan->elem_bytes;
    // image i: upload on s_up, filter in place on s_run, download on s_down; the upload of image i+1 and the
    // download of image i-1 run beside the kernels of image i (PCIe is full duplex)
    for (int i = 0; i < n; ++i) {
        const int b = i % NB;
        if (i >= NB) CUDA_TRY(cudaStreamWaitEvent(hp.s_up, hp.down[b], 0));      // buffer b is free again
        CUDA_TRY(cudaMemcpyAsync(hp.buf[b].p, in_host[i], bytes, cudaMemcpyHostToDevice, hp.s_up));
        CUDA_TRY(cudaEventRecord(hp.up[b], hp.s_up));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_run, hp.up[b], 0));
        if ((rc = rf_plan_execute(plan, hp.buf[b].p, hp.buf[b].p, hp.s_run))) return rc;
        CUDA_TRY(cudaEventRecord(hp.done[b], hp.s_run));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_down, hp.done[b], 0));
        CUDA_TRY(cudaMemcpyAsync(out_host[i], hp.buf[b].p, bytes, cudaMemcpyDeviceToHost, hp.s_down));
        CUDA_TRY(cudaEventRecord(hp.down[b], hp.s_down));
    }
    CUDA_TRY(cudaStreamSynchronize(hp.s_down));
    return rf_plan_check(plan);
}

int rf_plan_execute_host(rf_plan* plan, const void* in_host, void* out_host)
{
    return rf_plan_execute_host_batch(plan, 1, &in_host, &out_host);
}

int rf_stencil_execute(int ndim, const int64_t* extent, int dtype, int ntaps, const rf_tap* taps, float post_scale,
                       const void* in_dev, const void* in2_dev, void* out_dev, void* stream)
{
    if (ndim < 1 || ndim > RF_MAX_DIMS || !extent || !taps) return fail(RF_EINVAL, "bad stencil descriptor");
    if (ntaps < 1 || ntaps > RF_MAX_TAPS) return fail(RF_EINVAL, "a stencil has 1..%d taps", RF_MAX_TAPS);
    if (dtype != RF_F32 && dtype != RF_I32 && dtype != RF_U32) return fail(RF_EUNSUPPORTED, "stencils need a 32-bit element type");
    if (!in_dev || !out_dev) return fail(RF_EINVAL, "null buffer");
    if (in_dev == out_dev || in2_dev == out_dev) return fail(RF_EINVAL, "a stencil cannot run in place");
    for (int t = 0; t < ntaps; ++t)
        if (taps[t].source < 0 || taps[t].source > 1 || (taps[t].source == 1 && !in2_dev))
            return fail(RF_EINVAL, "tap %d reads a source that was not given", t);
    StencilParams p;
    std::memset(&p, 0, sizeof(p));
    p.ndim = ndim; p.ntaps = ntaps; p.total = 1; p.post_scale = post_scale;
    for (int d = 0; d < ndim; ++d) { if (extent[d] < 0) return fail(RF_EINVAL, "negative extent"); p.extent[d] = extent[d]; p.total *= extent[d]; }
    for (int t = 0; t < ntaps; ++t) p.tap[t] = taps[t];
    if (p.total == 0) return RF_OK;
    const unsigned blocks = (unsigned)std::min<int64_t>((p.total + 255) / 256, 148 * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (ndim <= 2 && p.extent[0] < 0x7fffff00LL && (ndim < 2 || p.extent[1] < 0x7fffff00LL)) {
        // about 16 blocks per SM, each a band of rows (RFB_STENCIL_BANDS overrides the band count)
        const int64_t gx = (p.extent[0] + 1023) / 1024, rows = ndim > 1 ? p.extent[1] : 1;
        static const int bands_env = getenv("RFB_STENCIL_BANDS") ? atoi(getenv("RFB_STENCIL_BANDS")) : 0;
        const int64_t want = bands_env > 0 ? bands_env : std::max<int64_t>(1, (148 * 16 + gx - 1) / gx);
        const dim3 grid((unsigned)gx, (unsigned)std::min<int64_t>(std::min<int64_t>(rows, want), 65535));
        if (dtype == RF_F32) stencil2d_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)in_dev, (const float*)in2_dev, (float*)out_dev);
        else                 stencil2d_kernel<uint32_t><<<grid, 256, 0, st>>>(p, (const uint32_t*)in_dev, (const uint32_t*)in2_dev, (uint32_t*)out_dev);
    } else if (dtype == RF_F32) stencil_kernel<float><<<blocks, 256, 0, st>>>(p, (const float*)in_dev, (const float*)in2_dev, (float*)out_dev);
    else                        stencil_kernel<uint32_t><<<blocks, 256, 0, st>>>(p, (const uint32_t*)in_dev, (const uint32_t*)in2_dev, (uint32_t*)out_dev);
    CUDA_TRY(cudaGetLastError());
    return RF_OK;
}

int rf_plan_profile(rf_plan* plan, const void* in_dev, void* out_dev, int iters, float* ms_per_iter)
{
    if (!plan || !ms_per_iter || iters < 1) return fail(RF_EINVAL, "bad argument");
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
    int rc = rf_plan_execute(plan, in_dev, out_dev, nullptr);          // warm-up (lib/recfilter.cpp:995-997)
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(a, 0));
    for (int i = 0; i < iters; ++i)
        if ((rc = rf_plan_execute(plan, in_dev, out_dev, nullptr))) return rc;
    CUDA_TRY(cudaEventRecord(b, 0));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_per_iter = ms / iters;
    return rf_plan_check(plan);
}

size_t rf_plan_shard_tail_bytes(const rf_plan* plan)
{
    if (!plan || plan->shard_pass < 0) return 0;
    return plan->passes[plan->shard_pass]->shard_tail_elems() * 4;
}

int rf_plan_stage1(rf_plan* plan, const void* in_dev, void* out_dev, void* tails_dev, void* stream)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (plan->shard_pass < 0) return fail(RF_EINVAL, "plan is not sharded");
    if (int drc = check_device(plan)) return drc;
    cudaStream_t st = (cudaStream_t)stream;
    const void* src = in_dev;
    for (int i = 0; i <= plan->shard_pass; ++i) {
        auto& p = plan->passes[i];
        int rc;
        if ((rc = p->run_tails(src, out_dev, st))) return rc;
        if (i == plan->shard_pass) return p->run_carries(nullptr, tails_dev, st, 1);
        if ((rc = p->run_carries(nullptr, nullptr, st))) return rc;
        if ((rc = p->run_final(src, out_dev, st))) return rc;
        src = out_dev;
    }
    return RF_OK;
}

int rf_plan_stage2(rf_plan* plan, const void* in_dev, void* out_dev, const void* gathered_tails_dev,
                   int nshards, int shard_rank, void* stream)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (plan->shard_pass < 0) return fail(RF_EINVAL, "plan is not sharded");
    if (nshards < 1 || shard_rank < 0 || shard_rank >= nshards) return fail(RF_EINVAL, "bad shard rank");
    if (int drc = check_device(plan)) return drc;
    cudaStream_t st = (cudaStream_t)stream;
    const void* src = plan->shard_pass == 0 ? in_dev : out_dev;
    for (int i = plan->shard_pass; i < (int)plan->passes.size(); ++i) {
        auto& p = plan->passes[i];
        int rc;
        if (i == plan->shard_pass) {
            if ((rc = p->shard_resolve(gathered_tails_dev, nshards, shard_rank, st))) return rc;
            // tails of K1 are still valid: redo the carry completion with the incoming shard carries
            if ((rc = p->run_carries(p->ext_buffer(), nullptr, st, 2))) return rc;
        } else {
            if ((rc = p->run_tails(src, out_dev, st))) return rc;
            if ((rc = p->run_carries(nullptr, nullptr, st))) return rc;
        }
        if ((rc = p->run_final(src, out_dev, st))) return rc;
        src = out_dev;
    }
    return RF_OK;
}

int rf_plan_shard_neighbors_suffice(const rf_plan* plan)
{
    if (!plan || plan->shard_pass < 0) return 0;
    return plan->passes[plan->shard_pass]->shard_neighbors_suffice() ? 1 : 0;
}

int rf_plan_shard_vectors(const rf_plan* plan)
{
    if (!plan || plan->shard_pass < 0) return 0;
    return plan->passes[plan->shard_pass]->shard_vectors();
}

int rf_plan_shard_resolve_lines(rf_plan* plan, const void* gathered_tails_dev, int nshards, int64_t nlines,
                                void* ext_all_dev, void* stream)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (plan->shard_pass < 0) return fail(RF_EINVAL, "plan is not sharded");
    if (nshards < 1 || nlines < 0 || !gathered_tails_dev || !ext_all_dev) return fail(RF_EINVAL, "bad argument");
    return plan->passes[plan->shard_pass]->shard_resolve_lines(gathered_tails_dev, nshards, nlines, ext_all_dev, (cudaStream_t)stream);
}

int rf_plan_stage2_ext(rf_plan* plan, const void* in_dev, void* out_dev, const void* ext_dev, void* stream)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    if (plan->shard_pass < 0) return fail(RF_EINVAL, "plan is not sharded");
    if (!ext_dev) return fail(RF_EINVAL, "null carries");
    if (int drc = check_device(plan)) return drc;
    cudaStream_t st = (cudaStream_t)stream;
    const void* src = plan->shard_pass == 0 ? in_dev : out_dev;
    for (int i = plan->shard_pass; i < (int)plan->passes.size(); ++i) {
        auto& p = plan->passes[i];
        int rc;
        if (i == plan->shard_pass) {
            if ((rc = p->run_carries(ext_dev, nullptr, st, 2))) return rc;
        } else {
            if ((rc = p->run_tails(src, out_dev, st))) return rc;
            if ((rc = p->run_carries(nullptr, nullptr, st))) return rc;
        }
        if ((rc = p->run_final(src, out_dev, st))) return rc;
        src = out_dev;
    }
    return RF_OK;
}

int rf_plan_check(rf_plan* plan)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    CUDA_TRY(cudaDeviceSynchronize());
    for (auto& p : plan->passes) {
        const uint32_t* flag = p->error_flag();
        if (!flag) continue;
        uint32_t v = 0;
        CUDA_TRY(cudaMemcpy(&v, flag, sizeof(v), cudaMemcpyDeviceToHost));
        if (v) return fail(RF_EINTERNAL, "a single-pass kernel gave up waiting for a predecessor tile (look-back spin limit)");
    }
    return RF_OK;
}

int rf_plan_stage_timing(rf_plan* plan, int enable)
{
    if (!plan) return fail(RF_EINVAL, "null plan");
    plan->timer.reset();
    plan->timer.on = enable != 0;
    return RF_OK;
}

int rf_plan_stage_times(rf_plan* plan, double* ms, long* counts, int n)
{
    if (!plan || !ms || !counts) return fail(RF_EINVAL, "null argument");
    plan->timer.collect();
    for (int i = 0; i < n && i < ST_COUNT; ++i) { ms[i] = plan->timer.ms[i]; counts[i] = plan->timer.count[i]; }
    return RF_OK;
}

int rf_clock_begin(void* stream, void** clock)
{
    if (!clock) return fail(RF_EINVAL, "null argument");
    cudaEvent_t* ev = new (std::nothrow) cudaEvent_t[2];
    if (!ev) return fail(RF_ENOMEM, "out of host memory");
    CUDA_TRY(cudaEventCreate(&ev[0])); CUDA_TRY(cudaEventCreate(&ev[1]));
    CUDA_TRY(cudaEventRecord(ev[0], (cudaStream_t)stream));
    *clock = ev;
    return RF_OK;
}

int rf_clock_end(void* clock, void* stream, float* ms)
{
    if (!clock || !ms) return fail(RF_EINVAL, "null argument");
    cudaEvent_t* ev = (cudaEvent_t*)clock;
    CUDA_TRY(cudaEventRecord(ev[1], (cudaStream_t)stream));
    CUDA_TRY(cudaEventSynchronize(ev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, ev[0], ev[1]));
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    delete[] ev;
    return RF_OK;
}

int rf_malloc(void** p, size_t bytes) { if (!p) return fail(RF_EINVAL, "null"); CUDA_TRY(cudaMalloc(p, bytes ? bytes : 1)); return RF_OK; }
int rf_free(void* p) { CUDA_TRY(cudaFree(p)); return RF_OK; }
int rf_memcpy_h2d(void* d, const void* s, size_t n) { CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)); return RF_OK; }
int rf_memcpy_d2h(void* d, const void* s, size_t n) { CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); return RF_OK; }
int rf_memset(void* d, int v, size_t n) { CUDA_TRY(cudaMemset(d, v, n)); return RF_OK; }
int rf_malloc_host(void** p, size_t bytes) { if (!p) return fail(RF_EINVAL, "null"); CUDA_TRY(cudaMallocHost(p, bytes ? bytes : 1)); return RF_OK; }
int rf_free_host(void* p) { CUDA_TRY(cudaFreeHost(p)); return RF_OK; }
int rf_synchronize(void) { CUDA_TRY(cudaDeviceSynchronize()); return RF_OK; }

} // extern "C"
#pragma GCC visibility pop
