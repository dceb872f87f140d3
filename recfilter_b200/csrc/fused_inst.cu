/*
 * fused_inst.cu -- explicit instantiation of the fused fast path for one padded filter
 * order (compile with -DRFB_R=<1|2|3|4>); one object per order so the orders build in
 * parallel.
 */
#include <type_traits>
#include <utility>
#include <cstring>
#include <cstdlib>
#include "fused.cuh"

#ifndef RFB_R
#error "compile with -DRFB_R=<order>"
#endif

namespace rfb {

// launch with programmatic stream serialization (PDL, see fused.cuh); RFB_NO_PDL=1 turns it off
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
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

template <typename CT, int R, int TS, bool RAGGED>
static cudaError_t launch_fused_tile_TS(const FusedParams<CT, R>& p, const void* in, void* out, int mode,
                                        cudaStream_t st)
{
    const int64_t nblocks = (int64_t)p.nbx * p.nbd * (p.No_launch > 0 ? p.No_launch : p.No);
    if (nblocks <= 0) return cudaSuccess;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = fused_tile_smem_bytes(TS, mode == FMODE_P2 ? fused_p2_carry_words(p.mx, p.md, R, TS, p.local, p.sdk) : 0);
    static bool attr_set_dev[RFB_MAX_DEVICES] = {};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    const bool cacheable = dev >= 0 && dev < RFB_MAX_DEVICES;
    if (!cacheable || !attr_set_dev[dev]) {
        cudaError_t e;
        e = cudaFuncSetAttribute(fused_tile_kernel<CT, R, TS, FMODE_P1, RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fused_tile_smem_bytes(TS));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fused_tile_kernel<CT, R, TS, FMODE_P2, RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)fused_tile_smem_bytes(TS, 2 * FMAX_SCANS * R * TS));
        if (e != cudaSuccess) return e;
        if (cacheable) attr_set_dev[dev] = true;
    }
    // L2 prefetch distance (blocks): about one wave of resident CTAs; RFB_PREFETCH overrides (0 = off)
    static const int prefetch_env = getenv("RFB_PREFETCH") ? atoi(getenv("RFB_PREFETCH")) : -1;
    FusedParams<CT, R> pp = p;
    // measured on C3: P1 (read only) gains ~5 % at a distance of 1-2 CTAs per SM, P2 (read + write) loses
    pp.prefetch = mode == FMODE_P1 ? (prefetch_env >= 0 ? prefetch_env : 2 * 148) : 0;
    const bool is_float = std::is_same<CT, float>::value;
    CUtensorMap tm_in, tm_out;
    cudaError_t e = make_tile_map(&tm_in, in, p.Nx, p.No * p.Nd, TS, is_float);
    if (e != cudaSuccess) return e;
    if (mode == FMODE_P1) {
        return launch_pdl(fused_tile_kernel<CT, R, TS, FMODE_P1, RAGGED>, dim3((unsigned)nblocks), dim3(TS), smem, st, pp, tm_in, tm_in);
    } else {
        e = make_tile_map(&tm_out, out, p.Nx, p.No * p.Nd, TS, is_float);
        if (e != cudaSuccess) return e;
        if constexpr (std::is_same<CT, float>::value && R <= 4) {
            if (p.epilogue) {
                static bool epi_attr_dev[RFB_MAX_DEVICES] = {};
                if (!cacheable || !epi_attr_dev[dev]) {
                    e = cudaFuncSetAttribute(fused_tile_kernel<CT, R, TS, FMODE_P2, RAGGED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)fused_tile_smem_bytes(TS, 2 * FMAX_SCANS * R * TS));
                    if (e != cudaSuccess) return e;
                    if (cacheable) epi_attr_dev[dev] = true;
                }
                return launch_pdl(fused_tile_kernel<CT, R, TS, FMODE_P2, RAGGED, true>, dim3((unsigned)nblocks), dim3(TS), smem, st, pp, tm_in, tm_out);
            }
        } else if (p.epilogue) return cudaErrorInvalidValue;
        return launch_pdl(fused_tile_kernel<CT, R, TS, FMODE_P2, RAGGED>, dim3((unsigned)nblocks), dim3(TS), smem, st, pp, tm_in, tm_out);
    }
}

template <typename CT, int R>
static cudaError_t launch_fused_tile_T(const FusedParams<CT, R>& p, const void* in, void* out, int mode, int ts,
                                       cudaStream_t st)
{
    const bool ragged = !p.signal && (p.Nx % ts != 0 || p.Nd % ts != 0);
    if (ts == 128) return ragged ? launch_fused_tile_TS<CT, R, 128, true>(p, in, out, mode, st)
                                 : launch_fused_tile_TS<CT, R, 128, false>(p, in, out, mode, st);
    if (ts == 64)  return ragged ? launch_fused_tile_TS<CT, R, 64, true>(p, in, out, mode, st)
                                 : launch_fused_tile_TS<CT, R, 64, false>(p, in, out, mode, st);
    return cudaErrorInvalidValue;
}

template <typename CT, int R, int S, int MAXT>
static cudaError_t launch_fchain_M(const FChainParams<CT, R>& p, unsigned grid, dim3 block, size_t smem, cudaStream_t st)
{
    // the opt-in is per device: remember the largest size set on each
    static size_t attr_bytes_dev[RFB_MAX_DEVICES] = {};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    const bool cacheable = dev >= 0 && dev < RFB_MAX_DEVICES;
    if (smem > 48u * 1024u && (!cacheable || smem > attr_bytes_dev[dev])) {
        cudaError_t e = cudaFuncSetAttribute(fchain_kernel<CT, R, S, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cacheable) attr_bytes_dev[dev] = smem;
    }
    return launch_pdl(fchain_kernel<CT, R, S, MAXT>, dim3(grid), block, smem, st, p);
}
template <typename CT, int R, int S>
static cudaError_t launch_fchain_S(const FChainParams<CT, R>& p, unsigned grid, dim3 block, size_t smem, cudaStream_t st)
{
    static const bool small_off = getenv("RFB_CHAIN_NO256") && atoi(getenv("RFB_CHAIN_NO256")) != 0;
    // measured on C3: with one wave of blocks (one 8192^2 image: 256 blocks) the register-rich instantiation is
    // faster (182 vs 190 us per image); the three-blocks-per-SM one pays off once there are several waves
    if (block.x * block.y <= 256 && grid >= 3 * 148 && !small_off) return launch_fchain_M<CT, R, S, 256>(p, grid, block, smem, st);
    return launch_fchain_M<CT, R, S, 512>(p, grid, block, smem, st);
}

template <typename CT, int R>
static cudaError_t launch_fchain_T(const FChainParams<CT, R>& pin, cudaStream_t st)
{
    FChainParams<CT, R> p = pin;
    if (p.l1 == 0) { p.l0 = 0; p.l1 = p.nl; }
    if (p.nl <= 0 || p.nb <= 0 || p.l1 <= p.l0) return cudaSuccess;
    if (p.l0 % 32 != 0 || p.l1 > p.nl) return cudaErrorInvalidConfiguration;
    if (p.nseg < 1 || p.nseg > 16 || p.S < 1 || p.S > FMAX_SCANS || p.L < 1) return cudaErrorInvalidConfiguration;
    const dim3 block(32, p.nseg);
    const unsigned grid = (unsigned)((p.l1 - p.l0 + 31) / 32);
    const size_t smem = fchain_smem_bytes(p.S, p.nseg, R, p.L, p.nb, p.A ? p.sdk : 0);
    if (smem > 227u * 1024u) return cudaErrorInvalidConfiguration;
    if ((int64_t)p.S * R * p.nb * p.nl > 0x7fffffffLL) return cudaErrorInvalidConfiguration;   // 32-bit offsets
    if (p.sJ <= 0 || p.sL <= 0) return cudaErrorInvalidConfiguration;
    if constexpr (R > 4) {                                   // high orders: one scan per dimension (register budget)
        if (p.S != 1) return cudaErrorInvalidConfiguration;
        return launch_fchain_S<CT, R, 1>(p, grid, block, smem, st);
    } else {
        switch (p.S) {
        case 1: return launch_fchain_S<CT, R, 1>(p, grid, block, smem, st);
        case 2: return launch_fchain_S<CT, R, 2>(p, grid, block, smem, st);
        case 3: return launch_fchain_S<CT, R, 3>(p, grid, block, smem, st);
        default: return launch_fchain_S<CT, R, 4>(p, grid, block, smem, st);
        }
    }
}

template <typename CT, int R>
static cudaError_t launch_flocal_T(const FLocalParams<CT, R>& p, cudaStream_t st)
{
    if (p.nl <= 0 || p.nb <= 0) return cudaSuccess;
    if (p.S < 1 || p.S > 2 || p.nb > 65535) return cudaErrorInvalidConfiguration;
    return launch_pdl(flocal_kernel<CT, R>, dim3((unsigned)((p.nl + 127) / 128), (unsigned)((p.nb + FLOCAL_TPT - 1) / FLOCAL_TPT)), dim3(128), 0, st, p);
}

template <typename CT, int R>
static cudaError_t launch_fcross_T(const FCrossParams<CT, R>& pin, int ts, cudaStream_t st)
{
    FCrossParams<CT, R> p = pin;
    if (p.w1 == 0) { p.w0 = 0; p.w1 = (int64_t)p.nbx * p.nbd * p.No; }
    const int64_t ntiles = p.w1 - p.w0;
    if (ntiles <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((ntiles + 3) / 4);
    if (p.local == 2) {                                      // tails mode: the x tails are corrected in place
        if (ts == 128) return launch_pdl(fcrossA_kernel<CT, R, 128, true>, dim3(grid), dim3(128), 0, st, p);
        if (ts == 64)  return launch_pdl(fcrossA_kernel<CT, R, 64, true>, dim3(grid), dim3(128), 0, st, p);
        return cudaErrorInvalidValue;
    }
    if (ts == 128) return launch_pdl(fcrossA_kernel<CT, R, 128, false>, dim3(grid), dim3(128), 0, st, p);
    if (ts == 64)  return launch_pdl(fcrossA_kernel<CT, R, 64, false>, dim3(grid), dim3(128), 0, st, p);
    return cudaErrorInvalidValue;
}

template <typename CT, int R>
static cudaError_t launch_fused_stream_T(const FStreamParams<CT, R>& sp, const void* in, void* out, cudaStream_t st)
{
    constexpr int TS = 128;
    const FusedParams<CT, R>& p = sp.t;
    const int64_t nblocks = (int64_t)(sp.rows + sp.lag_p) * sp.step;
    if (nblocks <= 0) return cudaSuccess;
    if (nblocks > 0x7fffffffLL || p.Nx % TS || p.Nd % TS) return cudaErrorInvalidConfiguration;
    const size_t smem = fused_tile_smem_bytes(TS, fused_p2_carry_words(p.mx, p.md, R, TS, 1, p.sdk));
    static size_t attr_bytes_dev[RFB_MAX_DEVICES] = {};
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    const bool cacheable = dev >= 0 && dev < RFB_MAX_DEVICES;
    if (!cacheable || smem > attr_bytes_dev[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fused_stream_kernel<CT, R, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (cacheable) attr_bytes_dev[dev] = smem;
    }
    const bool is_float = std::is_same<CT, float>::value;
    CUtensorMap tm_in, tm_out;
    cudaError_t e = make_tile_map(&tm_in, in, p.Nx, p.No * p.Nd, TS, is_float);
    if (e != cudaSuccess) return e;
    e = make_tile_map(&tm_out, out, p.Nx, p.No * p.Nd, TS, is_float);
    if (e != cudaSuccess) return e;
    fused_stream_kernel<CT, R, TS><<<dim3((unsigned)nblocks), dim3(TS), smem, st>>>(sp, tm_in, tm_out);
    return cudaGetLastError();
}

#define RFB_CAT_(a, b) a##b
#define RFB_CAT(a, b) RFB_CAT_(a, b)

cudaError_t RFB_CAT(launch_fused_tile_f, RFB_R)(const FusedParams<float, RFB_R>& p, const void* in, void* out, int mode, int ts, cudaStream_t st)
{ return launch_fused_tile_T<float, RFB_R>(p, in, out, mode, ts, st); }
cudaError_t RFB_CAT(launch_fused_tile_u, RFB_R)(const FusedParams<uint32_t, RFB_R>& p, const void* in, void* out, int mode, int ts, cudaStream_t st)
{ return launch_fused_tile_T<uint32_t, RFB_R>(p, in, out, mode, ts, st); }
#if RFB_R <= 4
cudaError_t RFB_CAT(launch_fused_stream_f, RFB_R)(const FStreamParams<float, RFB_R>& p, const void* in, void* out, cudaStream_t st)
{ return launch_fused_stream_T<float, RFB_R>(p, in, out, st); }
#endif
cudaError_t RFB_CAT(launch_fchain_f, RFB_R)(const FChainParams<float, RFB_R>& p, cudaStream_t st)
{ return launch_fchain_T<float, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_fchain_u, RFB_R)(const FChainParams<uint32_t, RFB_R>& p, cudaStream_t st)
{ return launch_fchain_T<uint32_t, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_flocal_f, RFB_R)(const FLocalParams<float, RFB_R>& p, cudaStream_t st)
{ return launch_flocal_T<float, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_flocal_u, RFB_R)(const FLocalParams<uint32_t, RFB_R>& p, cudaStream_t st)
{ return launch_flocal_T<uint32_t, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_fcross_f, RFB_R)(const FCrossParams<float, RFB_R>& p, int ts, cudaStream_t st)
{ return launch_fcross_T<float, RFB_R>(p, ts, st); }
cudaError_t RFB_CAT(launch_fcross_u, RFB_R)(const FCrossParams<uint32_t, RFB_R>& p, int ts, cudaStream_t st)
{ return launch_fcross_T<uint32_t, RFB_R>(p, ts, st); }

} // namespace rfb
