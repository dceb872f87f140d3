// bw_probe2.cu -- development probe (not part of the product).
//   E1  read-only / copy bandwidth against the working-set size (is 268 MB long enough to reach the asymptote?)
//   E2  cross-kernel L2 reuse: kernel A reads X MB, kernel B reads the same X MB again
//   E3  the two-pass pattern of the filter: A reads 268 MB forwards (tail of it marked evict_last),
//       B reads it backwards and writes 268 MB (streaming hints) -- how much of the re-read can L2 serve?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe2 bw_probe2.cu && ./bw_probe2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t pol_last()  { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ float4 ldh(const float4* p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void sth(float4* p, float4 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

constexpr int U = 4;
// chunked sweep: chunk c covers float4 indices [c*CH, (c+1)*CH), CH = blockDim*U; chunks are dealt round-robin to CTAs
// hint: 0 none, 1 = loads evict_first below keep_from / evict_last at or above it; stores evict_first
__global__ void __launch_bounds__(256) sweep(const float4* __restrict__ a, float4* __restrict__ b, size_t n4, int rev, int hint,
                                            size_t keep_from4, int store)
{
    const size_t CH = (size_t)blockDim.x * U;
    const size_t nch = n4 / CH;
    const uint64_t pl = pol_last(), pf = pol_first();
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t c = blockIdx.x; c < nch; c += gridDim.x) {
        const size_t cc = rev ? nch - 1 - c : c;
        const size_t base = cc * CH + threadIdx.x;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + (size_t)u * blockDim.x;
            v[u] = hint ? ldh(a + i, i >= keep_from4 ? pl : pf) : __ldg(a + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + (size_t)u * blockDim.x;
            if (store) { if (hint) sth(b + i, v[u], pf); else b[i] = v[u]; }
            else { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    if (!store && acc.x + acc.y + acc.z + acc.w == 123.456f) b[0] = acc;
}
__global__ void nop() {}

int main()
{
    const size_t MAXB = (size_t)4 << 30;
    float *a, *b, *flush;
    CK(cudaMalloc(&a, MAXB)); CK(cudaMalloc(&b, MAXB)); CK(cudaMalloc(&flush, 512u << 20));
    CK(cudaMemset(a, 1, MAXB)); CK(cudaMemset(b, 0, MAXB));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int G = sms * 8;

    // launch overhead seen by events
    {
        float best = 1e9f;
        for (int it = 0; it < 10; ++it) {
            CK(cudaEventRecord(e0)); nop<<<1, 32>>>(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("E0 empty kernel between events: %.2f us\n", best * 1e3);
    }
    // E1
    for (size_t mb : {64, 128, 256, 512, 1024, 2048, 4096}) {
        const size_t n4 = mb * (1u << 20) / 16;
        for (int store = 0; store < 2; ++store) {
            for (int g : {sms * 4, sms * 8, sms * 16}) {
                float best = 1e9f;
                for (int it = 0; it < 4; ++it) {
                    CK(cudaMemsetAsync(flush, it, 512u << 20));
                    CK(cudaEventRecord(e0)); sweep<<<g, 256>>>((const float4*)a, (float4*)b, n4, 0, 0, 0, store);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
                }
                printf("E1 %-5s %5zu MB grid %4d: %8.1f us  %6.0f GB/s\n", store ? "copy" : "read", mb, g, best * 1e3,
                       (double)mb * 1.048576 * (1 + store) / best);
            }
        }
    }
    // E2
    for (int hint = 0; hint < 2; ++hint)
        for (size_t mb : {16, 32, 48, 64, 80, 96, 112, 128, 160}) {
            const size_t n4 = mb * (1u << 20) / 16;
            float best = 1e9f;
            for (int it = 0; it < 4; ++it) {
                CK(cudaMemsetAsync(flush, it, 512u << 20));
                sweep<<<G, 256>>>((const float4*)a, (float4*)b, n4, 0, hint, 0, 0);
                CK(cudaEventRecord(e0)); sweep<<<G, 256>>>((const float4*)a, (float4*)b, n4, 0, hint, 0, 0);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            printf("E2 re-read %4zu MB (%s): %8.1f us  %6.0f GB/s\n", mb, hint ? "evict_last" : "no hint", best * 1e3,
                   (double)mb * 1.048576 / best);
        }
    // E3
    {
        const size_t mb = 256, n4 = mb * (1u << 20) / 16;
        for (int rev = 0; rev < 2; ++rev)
            for (int keep : {-1, 0, 32, 48, 64, 80, 96, 112}) {
                const int hint = keep >= 0;
                const size_t keep_from4 = keep > 0 ? n4 - (size_t)keep * (1u << 20) / 16 : n4;
                float bestA = 1e9f, bestB = 1e9f;
                for (int it = 0; it < 4; ++it) {
                    CK(cudaMemsetAsync(flush, it, 512u << 20));
                    CK(cudaEventRecord(e0));
                    sweep<<<G, 256>>>((const float4*)a, (float4*)b, n4, 0, hint, keep_from4, 0);
                    CK(cudaEventRecord(e1));
                    sweep<<<G, 256>>>((const float4*)a, (float4*)b, n4, rev, hint, n4, 1);
                    CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
                    float ma, mb2; CK(cudaEventElapsedTime(&ma, e0, e1)); CK(cudaEventElapsedTime(&mb2, e1, e2));
                    if (ma + mb2 < bestA + bestB) { bestA = ma; bestB = mb2; }
                }
                printf("E3 pass B %s, keep %4d MB evict_last (%s): A %7.1f us  B %7.1f us  total %7.1f us\n", rev ? "reversed" : "forward ",
                       keep, hint ? "hints" : "no hints", bestA * 1e3, bestB * 1e3, (bestA + bestB) * 1e3);
            }
    }
    return 0;
}
