// bw_probe.cu -- what HBM bandwidth does a tiled TMA access pattern reach on this part?
// Development probe (not part of the product): persistent CTAs stream TH x TW tiles of an
// 8192 x 8192 fp32 image through a ring of shared-memory buffers with TMA loads (and optionally
// TMA stores of the same tile to a second image), no arithmetic.  Compared with a plain
// grid-stride float4 copy.   nvcc -arch=sm_100a -O3 -o bw_probe bw_probe.cu && ./bw_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 :: "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma_ld(void* dst, const void* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(s32(dst)), "l"(map), "r"(x), "r"(y), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_st(const void* map, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(map), "r"(x), "r"(y), "r"(s32(src)) : "memory");
}

struct Cfg { int W, H, TW, TH, BW, nbuf, store, order; };

// one producer thread per CTA does everything (loads, waits, stores): pure memory-system probe
__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, Cfg c)
{
    extern __shared__ __align__(16) unsigned char raw[];
    unsigned char* base = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
    const int tile_bytes = c.TW * c.TH * 4;
    uint64_t* full = (uint64_t*)(base + c.nbuf * tile_bytes);
    if (threadIdx.x != 0) return;
    for (int i = 0; i < c.nbuf; ++i) mbar_init(full + i, 1);
    const int ntx = c.W / c.TW, nty = c.H / c.TH;
    const int64_t nt = (int64_t)ntx * nty;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t mine = first < nt ? (nt - first + stride - 1) / stride : 0;
    const int nbox = c.TW / c.BW;
    auto coords = [&](int64_t k, int& x, int& y) {
        int64_t t = first + k * stride;
        if (c.order == 1) t = nt - 1 - t;
        x = (int)(t % ntx) * c.TW; y = (int)(t / ntx) * c.TH;
    };
    for (int64_t k = 0; k < mine + c.nbuf; ++k) {
        const int b = (int)(k % c.nbuf);
        unsigned char* buf = base + b * tile_bytes;
        if (k >= c.nbuf) {
            mbar_wait(full + b, (uint32_t)(((k - c.nbuf) / c.nbuf) & 1));
            if (c.store) {
                int x, y; coords(k - c.nbuf, x, y);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                for (int bb = 0; bb < nbox; ++bb) tma_st(&tout, x + bb * c.BW, y, buf + bb * (c.BW * c.TH * 4));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
        if (k < mine) {
            int x, y; coords(k, x, y);
            mbar_expect(full + b, tile_bytes);
            for (int bb = 0; bb < nbox; ++bb) tma_ld(buf + bb * (c.BW * c.TH * 4), &tin, x + bb * c.BW, y, full + b);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void copy4(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int store)
{
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = a[i];
        if (store) b[i] = v; else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    if (!store && acc.x + acc.y + acc.z + acc.w == 123.456f) b[0] = acc;
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int W = 8192, H = 8192;
    const size_t n = (size_t)W * H;
    float *a, *b, *flush;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&flush, 512u << 20));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(b, 0, n * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    enc_fn enc = (enc_fn)fn;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    auto time_it = [&](auto&& launch, double bytes, const char* name) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            CK(cudaMemsetAsync(flush, it, 512u << 20));          // flush L2
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-64s %8.1f us  %7.0f GB/s\n", name, best * 1e3, bytes / best / 1e6);
    };
    time_it([&] { copy4<<<sms * 16, 512>>>((const float4*)a, (float4*)b, n / 4, 0); }, n * 4.0, "plain float4 read");
    time_it([&] { copy4<<<sms * 16, 512>>>((const float4*)a, (float4*)b, n / 4, 1); }, n * 8.0, "plain float4 copy (R+W)");

    struct V { int TW, TH, BW, swz, nbuf, cps, store, order; };
    std::vector<V> vs = {
        {128, 128, 32, 1, 3, 1, 0, 0}, {128, 128, 32, 1, 3, 1, 1, 0},
        {128, 128, 32, 1, 2, 1, 0, 0}, {128, 128, 32, 1, 1, 3, 0, 0}, {128, 128, 32, 1, 1, 3, 1, 0},
        {128, 128, 128, 0, 3, 1, 0, 0}, {128, 128, 128, 0, 3, 1, 1, 0},
        {256, 64, 32, 1, 3, 1, 0, 0}, {256, 64, 256, 0, 3, 1, 0, 0}, {256, 64, 256, 0, 3, 1, 1, 0},
        {64, 64, 32, 1, 4, 3, 0, 0}, {64, 64, 32, 1, 4, 3, 1, 0}, {64, 64, 32, 1, 6, 2, 0, 0},
        {128, 64, 32, 1, 3, 2, 0, 0}, {128, 64, 32, 1, 3, 2, 1, 0},
        {128, 128, 32, 1, 3, 1, 1, 1},
        {256, 128, 32, 1, 1, 1, 0, 0}, {512, 32, 32, 1, 3, 1, 0, 0}, {1024, 16, 32, 1, 3, 1, 0, 0},
    };
    for (auto& v : vs) {
        CUtensorMap ti, to;
        cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)W * 4};
        cuuint32_t box[2] = {(cuuint32_t)v.BW, (cuuint32_t)v.TH}, es[2] = {1, 1};
        auto sw = v.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        if (enc(&ti, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
            enc(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, b, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); continue; }
        Cfg c{W, H, v.TW, v.TH, v.BW, v.nbuf, v.store, v.order};
        size_t smem = (size_t)v.nbuf * v.TW * v.TH * 4 + 1024 + 128;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        char name[160];
        snprintf(name, sizeof(name), "TMA tile %4dx%-3d box %3d swz%d ring %d, %d CTA/SM %s%s", v.TW, v.TH, v.BW, v.swz, v.nbuf, v.cps,
                 v.store ? "load+store" : "load only", v.order ? " reversed" : "");
        time_it([&] { probe<<<sms * v.cps, 32, smem>>>(ti, to, c); }, n * 4.0 * (1 + v.store), name);
    }
    return 0;
}
