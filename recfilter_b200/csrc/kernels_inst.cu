/*
 * kernels_inst.cu -- explicit instantiation of the kernel family for one padded
 * filter order (compile with -DRFB_R=<1|2|3|4|8|16|32>); one object per order so the
 * orders build in parallel.
 */
#include <cstdlib>
#include "kernels.cuh"

#ifndef RFB_R
#error "compile with -DRFB_R=<order>"
#endif

namespace rfb {

template <typename CT, int R>
static cudaError_t launch_tile_T(const PassParams<CT, R>& p, const void* in, void* out, int mode, cudaStream_t st)
{
    int64_t nblocks;
    if (p.signal_mode) nblocks = ((int64_t)(p.nbx + TILE - 1) / TILE) * p.nlx;
    else               nblocks = (int64_t)p.nbx * p.nbd * p.No;
    if (nblocks <= 0) return cudaSuccess;
    if (nblocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (mode == MODE_TAILS)
        tile_kernel<CT, R, MODE_TAILS><<<(unsigned)nblocks, TILE, 0, st>>>(p, (const CT*)in, (CT*)out);
    else
        tile_kernel<CT, R, MODE_FINAL><<<(unsigned)nblocks, TILE, 0, st>>>(p, (const CT*)in, (CT*)out);
    return cudaGetLastError();
}

template <typename CT, int R>
static cudaError_t launch_chain_T(const ChainParams<CT, R>& p, cudaStream_t st)
{
    const int64_t n1 = p.nl * p.nseg;
    if (n1 <= 0) return cudaSuccess;
    if constexpr (R >= 16) {
        // high orders: one warp per (segment, line), see kernels.cuh (RFB_NO_WARP_CHAIN=1: the thread-per-line kernels)
        static const bool off = getenv("RFB_NO_WARP_CHAIN") && atoi(getenv("RFB_NO_WARP_CHAIN")) != 0;
        if (!off) {
            chain_local_wkernel<CT, R><<<(unsigned)((n1 + 3) / 4), 128, 0, st>>>(p);
            if (p.nseg > 1) {
                chain_top_wkernel<CT, R><<<(unsigned)((p.nl + 3) / 4), 128, 0, st>>>(p);
                const int64_t n3 = p.nl * (p.nseg - 1);
                chain_fix_wkernel<CT, R><<<(unsigned)((n3 + 3) / 4), 128, 0, st>>>(p);
            }
            return cudaGetLastError();
        }
    }
    chain_local_kernel<CT, R><<<(unsigned)((n1 + 127) / 128), 128, 0, st>>>(p);
    if (p.nseg > 1) {
        chain_top_kernel<CT, R><<<(unsigned)((p.nl + 127) / 128), 128, 0, st>>>(p);
        const int64_t n3 = p.nl * (p.nseg - 1);
        chain_fix_kernel<CT, R><<<(unsigned)((n3 + 127) / 128), 128, 0, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename CT, int R>
static cudaError_t launch_cross_T(const CrossParams<CT, R>& p, cudaStream_t st)
{
    const int64_t nblocks = (int64_t)p.nbx * p.nbd * p.No;
    if (nblocks <= 0) return cudaSuccess;
    const size_t smem = ((size_t)p.md * p.mx * R * R + TILE / 32) * sizeof(typename TabType<CT>::type);
    cross_kernel<CT, R><<<(unsigned)nblocks, TILE, smem, st>>>(p);
    return cudaGetLastError();
}

#define RFB_CAT_(a, b) a##b
#define RFB_CAT(a, b) RFB_CAT_(a, b)

cudaError_t RFB_CAT(launch_tile_f, RFB_R)(const PassParams<float, RFB_R>& p, const void* in, void* out, int mode, cudaStream_t st)
{ return launch_tile_T<float, RFB_R>(p, in, out, mode, st); }
cudaError_t RFB_CAT(launch_tile_u, RFB_R)(const PassParams<uint32_t, RFB_R>& p, const void* in, void* out, int mode, cudaStream_t st)
{ return launch_tile_T<uint32_t, RFB_R>(p, in, out, mode, st); }
cudaError_t RFB_CAT(launch_chain_f, RFB_R)(const ChainParams<float, RFB_R>& p, cudaStream_t st)
{ return launch_chain_T<float, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_chain_u, RFB_R)(const ChainParams<uint32_t, RFB_R>& p, cudaStream_t st)
{ return launch_chain_T<uint32_t, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_cross_f, RFB_R)(const CrossParams<float, RFB_R>& p, cudaStream_t st)
{ return launch_cross_T<float, RFB_R>(p, st); }
cudaError_t RFB_CAT(launch_cross_u, RFB_R)(const CrossParams<uint32_t, RFB_R>& p, cudaStream_t st)
{ return launch_cross_T<uint32_t, RFB_R>(p, st); }

} // namespace rfb
