/*
 * lookback_inst.cu -- explicit instantiation of the single-pass look-back kernels for one padded
 * filter order (compile with -DRFB_R=<1|2|3|4|8>); one object per order so the orders build in parallel.
 */
#include <type_traits>
#include <utility>
#include <cstring>
#include <cstdlib>
#include "lookback.cuh"

#ifndef RFB_R
#error "compile with -DRFB_R=<order>"
#endif

namespace rfb {

template <typename... KArgs, typename... Args>
static cudaError_t lb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    static const bool use_pdl = !(getenv("RFB_NO_PDL") && atoi(getenv("RFB_NO_PDL")) != 0);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember it per device
template <typename K>
static cudaError_t lb_ensure_smem(K kernel, size_t bytes, bool (&done)[RFB_MAX_DEVICES])
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= RFB_MAX_DEVICES || !done[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < RFB_MAX_DEVICES) done[dev] = true;
    }
    return cudaSuccess;
}

template <typename CT, int R, int TS>
static cudaError_t launch_lb_tile_TS(const LBTileParams<CT, R>& p, const void* in, void* out, cudaStream_t st)
{
    const int64_t nblocks = (int64_t)p.nbx * p.nbd * p.No;
    if (nblocks <= 0) return cudaSuccess;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = lb_tile_smem_bytes(TS);
    static bool done[RFB_MAX_DEVICES] = {};
    cudaError_t e = lb_ensure_smem(lb_tile_kernel<CT, R, TS>, smem, done);
    if (e != cudaSuccess) return e;
    const bool is_float = std::is_same<CT, float>::value;
    CUtensorMap tm_in, tm_out;
    e = make_tile_map(&tm_in, in, p.Nx, p.No * p.Nd, TS, is_float);
    if (e != cudaSuccess) return e;
    e = make_tile_map(&tm_out, out, p.Nx, p.No * p.Nd, TS, is_float);
    if (e != cudaSuccess) return e;
    return lb_launch_pdl(lb_tile_kernel<CT, R, TS>, dim3((unsigned)nblocks), dim3(TS), smem, st, p, tm_in, tm_out);
}

template <typename CT, int R>
static cudaError_t launch_lb_tile_T(const LBTileParams<CT, R>& p, const void* in, void* out, int ts, cudaStream_t st)
{
    if constexpr (R > 4) {
        return cudaErrorInvalidValue;                        // the 2-D kernel is built for orders <= 4
    } else {
        if (ts == 128) return launch_lb_tile_TS<CT, R, 128>(p, in, out, st);
        if (ts == 64)  return launch_lb_tile_TS<CT, R, 64>(p, in, out, st);
        return cudaErrorInvalidValue;
    }
}

template <typename CT, int R, int NW>
static cudaError_t launch_lb_signal_NW(const LBSignalParams<CT, R>& p, const void* in, void* out, cudaStream_t st)
{
    constexpr int ROWS = 32 * NW;
    const int64_t nblocks = p.rows / ROWS;
    if (nblocks <= 0) return cudaSuccess;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)ROWS * 512 + 1024 + 64;
    static bool done[RFB_MAX_DEVICES] = {};
    cudaError_t e = lb_ensure_smem(lb_signal_kernel<CT, R, NW>, smem, done);
    if (e != cudaSuccess) return e;
    const bool is_float = std::is_same<CT, float>::value;
    CUtensorMap tm_in, tm_out;
    e = make_tile_map(&tm_in, in, 128, p.rows, ROWS, is_float);
    if (e != cudaSuccess) return e;
    e = make_tile_map(&tm_out, out, 128, p.rows, ROWS, is_float);
    if (e != cudaSuccess) return e;
    return lb_launch_pdl(lb_signal_kernel<CT, R, NW>, dim3((unsigned)nblocks), dim3(ROWS), smem, st, p, tm_in, tm_out);
}
template <typename CT, int R>
static cudaError_t launch_lb_signal_T(const LBSignalParams<CT, R>& p, const void* in, void* out, cudaStream_t st)
{
    switch (p.tile_rows) {
    case 128: return launch_lb_signal_NW<CT, R, 4>(p, in, out, st);
    case 64:  return launch_lb_signal_NW<CT, R, 2>(p, in, out, st);
    case 32:  return launch_lb_signal_NW<CT, R, 1>(p, in, out, st);
    }
    return cudaErrorInvalidValue;
}

#define RFB_CAT_(a, b) a##b
#define RFB_CAT(a, b) RFB_CAT_(a, b)

cudaError_t RFB_CAT(launch_lb_tile_f, RFB_R)(const LBTileParams<float, RFB_R>& p, const void* in, void* out, int ts, cudaStream_t st)
{ return launch_lb_tile_T<float, RFB_R>(p, in, out, ts, st); }
cudaError_t RFB_CAT(launch_lb_tile_u, RFB_R)(const LBTileParams<uint32_t, RFB_R>& p, const void* in, void* out, int ts, cudaStream_t st)
{ return launch_lb_tile_T<uint32_t, RFB_R>(p, in, out, ts, st); }
cudaError_t RFB_CAT(launch_lb_signal_f, RFB_R)(const LBSignalParams<float, RFB_R>& p, const void* in, void* out, cudaStream_t st)
{ return launch_lb_signal_T<float, RFB_R>(p, in, out, st); }
cudaError_t RFB_CAT(launch_lb_signal_u, RFB_R)(const LBSignalParams<uint32_t, RFB_R>& p, const void* in, void* out, cudaStream_t st)
{ return launch_lb_signal_T<uint32_t, RFB_R>(p, in, out, st); }

} // namespace rfb
