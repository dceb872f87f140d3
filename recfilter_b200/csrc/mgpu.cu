/*
 * mgpu.cu -- one filter over several GPUs of ONE process (C ABI: rf_mgpu_*, include/recfilter_b200.h).
 *
 * The reference is single-GPU; this is the multi-GPU layer of SURVEY 8e for hosts that are one process -- the
 * C++ operator surface (RecFilter::realize / profile with RECFILTER_GPUS=n) sits on it.  The planner picks:
 *
 *   batch sharding   the outermost dimension carries no scans (a stack of images, the channels of apps/audio):
 *                    every GPU filters extent / n whole units with its own plan, nothing is exchanged;
 *   strip sharding   the outermost dimension is scanned: it is cut into n strips (rows of an image, z slabs of a
 *                    volume).  Every GPU runs stage 1 on its strip (rf_plan_stage1), PULLS the order-r boundary
 *                    tails of all strips from its peers (cudaMemcpyPeerAsync over NVLink, ordered by events
 *                    between the per-device streams -- no host copy, no collective library), and finishes with
 *                    stage 2 (rf_plan_stage2).  No image data crosses the link.
 *
 * One stream per device; the host thread only enqueues.
 */
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../../include/recfilter_b200.h"

namespace {

thread_local char m_err[640] = "";
int mfail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(m_err, sizeof(m_err), fmt, ap);
    va_end(ap);
    return code;
}
#define MCUDA_TRY(expr)                                                                  \
    do { cudaError_t e__ = (expr);                                                       \
         if (e__ != cudaSuccess)                                                         \
             return mfail(RF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define MRF_TRY(expr)                                                                    \
    do { int rc__ = (expr);                                                              \
         if (rc__ != RF_OK) return mfail(rc__, "%s: %s", #expr, rf_last_error());        \
    } while (0)

size_t elem_bytes(int dt)
{
    switch (dt) {
    case RF_F32: case RF_I32: case RF_U32: return 4;
    case RF_I16: case RF_U16: return 2;
    case RF_I8:  case RF_U8:  return 1;
    }
    return 0;
}

} // namespace

struct rf_mgpu {
    int n = 1;
    bool strips = false;                 // strip sharding (tail exchange) or batch sharding
    int dim = 0;                         // the dimension that is cut (the outermost one)
    rf_desc desc;                        // the whole filter
    size_t part_bytes = 0, tail_bytes = 0;
    std::vector<rf_plan*> plan;
    std::vector<void*> in, out, tails, gathered;
    std::vector<cudaStream_t> st;
    std::vector<cudaEvent_t> ev_stage1, ev_pulled, ev_t0, ev_t1;
    std::string text;
    int home = 0;                        // device that was current when the object was made
};

extern "C" {
#pragma GCC visibility push(default)

const char* rf_mgpu_last_error(void) { return m_err; }

void rf_mgpu_destroy(rf_mgpu* m)
{
    if (!m) return;
    for (int d = 0; d < m->n; ++d) {
        cudaSetDevice(d);
        if (d < (int)m->st.size() && m->st[d]) cudaStreamSynchronize(m->st[d]);
        if (d < (int)m->plan.size() && m->plan[d]) rf_plan_destroy(m->plan[d]);
        for (auto* v : { &m->in, &m->out, &m->tails, &m->gathered })
            if (d < (int)v->size() && (*v)[d]) cudaFree((*v)[d]);
        for (auto* v : { &m->ev_stage1, &m->ev_pulled, &m->ev_t0, &m->ev_t1 })
            if (d < (int)v->size() && (*v)[d]) cudaEventDestroy((*v)[d]);
        if (d < (int)m->st.size() && m->st[d]) cudaStreamDestroy(m->st[d]);
    }
    cudaSetDevice(m->home);
    delete m;
}

int rf_mgpu_create(const rf_desc* desc, int ngpus, rf_mgpu** out)
{
    if (!desc || !out) return mfail(RF_EINVAL, "null argument");
    *out = nullptr;
    const int have = rf_device_count();
    if (have < 1) return mfail(RF_ENODEVICE, "no CUDA device available: this engine has no CPU fallback");
    if (ngpus < 1 || ngpus > have) return mfail(RF_EINVAL, "%d GPUs requested, %d present", ngpus, have);
    if (desc->ndim < 1 || desc->ndim > RF_MAX_DIMS) return mfail(RF_EINVAL, "ndim must be 1..%d", RF_MAX_DIMS);
    const size_t eb = elem_bytes(desc->dtype);
    if (!eb) return mfail(RF_EINVAL, "unknown dtype %d", desc->dtype);
    const int D = desc->ndim - 1;
    bool scanned = false;
    for (int s = 0; s < desc->nscans; ++s) scanned = scanned || desc->scans[s].dim == D;
    if (ngpus > 1) {
        if (D == 0 && scanned) return mfail(RF_EUNSUPPORTED, "a 1-D filter cannot be cut: batch several signals instead");
        if (desc->extent[D] % ngpus) return mfail(RF_EUNSUPPORTED, "extent %lld of the outermost dimension is not divisible by %d GPUs",
                                                   (long long)desc->extent[D], ngpus);
        if (scanned && eb != 4) return mfail(RF_EUNSUPPORTED, "strip sharding needs a 32-bit element type");
    }
    rf_mgpu* m = new (std::nothrow) rf_mgpu();
    if (!m) return mfail(RF_ENOMEM, "out of host memory");
    m->n = ngpus; m->dim = D; m->strips = scanned && ngpus > 1; m->desc = *desc;
    cudaGetDevice(&m->home);
    int64_t inner = 1;
    for (int d = 0; d < D; ++d) inner *= desc->extent[d];
    const int64_t part = desc->extent[D] / ngpus;
    m->part_bytes = (size_t)(inner * part) * eb;
    m->plan.assign(ngpus, nullptr);
    m->in.assign(ngpus, nullptr); m->out.assign(ngpus, nullptr); m->tails.assign(ngpus, nullptr); m->gathered.assign(ngpus, nullptr);
    m->st.assign(ngpus, nullptr);
    m->ev_stage1.assign(ngpus, nullptr); m->ev_pulled.assign(ngpus, nullptr); m->ev_t0.assign(ngpus, nullptr); m->ev_t1.assign(ngpus, nullptr);
    int rc = RF_OK;
    for (int d = 0; d < ngpus && rc == RF_OK; ++d) {
        cudaError_t e = cudaSetDevice(d);
        if (e != cudaSuccess) { rc = mfail(RF_ECUDA, "cudaSetDevice(%d): %s", d, cudaGetErrorString(e)); break; }
        for (int p = 0; p < ngpus; ++p) {                    // direct NVLink copies where the topology allows
            if (p == d) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, p) == cudaSuccess && can) {
                e = cudaDeviceEnablePeerAccess(p, 0);
                if (e != cudaSuccess) cudaGetLastError();     // already enabled, or not possible: the copies still work
            }
        }
        rf_desc dd = *desc;
        dd.extent[D] = part;
        if (m->strips) { dd.opt.shard_dim = D; dd.opt.open_lo = d > 0; dd.opt.open_hi = d < ngpus - 1; }
        rc = rf_plan_create(&dd, &m->plan[d]);
        if (rc != RF_OK) { mfail(rc, "rf_plan_create on device %d: %s", d, rf_last_error()); break; }
        if ((e = cudaStreamCreateWithFlags(&m->st[d], cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaMalloc(&m->in[d], m->part_bytes ? m->part_bytes : 1)) != cudaSuccess ||
            (e = cudaMalloc(&m->out[d], m->part_bytes ? m->part_bytes : 1)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&m->ev_stage1[d], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&m->ev_pulled[d], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreate(&m->ev_t0[d])) != cudaSuccess || (e = cudaEventCreate(&m->ev_t1[d])) != cudaSuccess) {
            rc = mfail(RF_ECUDA, "device %d set-up: %s", d, cudaGetErrorString(e)); break;
        }
        if (m->strips) {
            m->tail_bytes = rf_plan_shard_tail_bytes(m->plan[d]);
            if ((e = cudaMalloc(&m->tails[d], m->tail_bytes ? m->tail_bytes : 1)) != cudaSuccess ||
                (e = cudaMalloc(&m->gathered[d], m->tail_bytes ? m->tail_bytes * ngpus : 1)) != cudaSuccess) {
                rc = mfail(RF_ECUDA, "device %d set-up: %s", d, cudaGetErrorString(e)); break;
            }
        }
    }
    if (rc != RF_OK) { rf_mgpu_destroy(m); return rc; }
    char head[512], plan0[4096] = "";
    rf_plan_describe(m->plan[0], plan0, sizeof(plan0));
    snprintf(head, sizeof(head), "recfilter_b200 multi-GPU plan: %d GPUs, dimension %d cut into %s of %lld%s\n  per GPU: ", ngpus, D,
             m->strips ? "strips" : "independent parts", (long long)part,
             m->strips ? " (order-r boundary tails pulled from the peers over NVLink, one exchange per call)" : " (nothing exchanged)");
    m->text = std::string(head) + plan0;
    cudaSetDevice(m->home);
    *out = m;
    return RF_OK;
}

int rf_mgpu_ngpus(const rf_mgpu* m) { return m ? m->n : 0; }

int rf_mgpu_describe(const rf_mgpu* m, char* buf, size_t n)
{
    if (!m || !buf || n == 0) return mfail(RF_EINVAL, "null argument");
    snprintf(buf, n, "%s", m->text.c_str());
    return RF_OK;
}

// one pass over the device-resident parts: in[d] -> out[d] on every device (asynchronous)
static int mgpu_run(rf_mgpu* m)
{
    const int n = m->n;
    if (!m->strips) {
        for (int d = 0; d < n; ++d) {
            MCUDA_TRY(cudaSetDevice(d));
            MRF_TRY(rf_plan_execute(m->plan[d], m->in[d], m->out[d], m->st[d]));
        }
        return RF_OK;
    }
    for (int d = 0; d < n; ++d) {
        MCUDA_TRY(cudaSetDevice(d));
        for (int p = 0; p < n; ++p)                           // nobody may still be pulling the tails of the call before
            if (p != d) MCUDA_TRY(cudaStreamWaitEvent(m->st[d], m->ev_pulled[p], 0));
        MRF_TRY(rf_plan_stage1(m->plan[d], m->in[d], m->out[d], m->tails[d], m->st[d]));
        MCUDA_TRY(cudaEventRecord(m->ev_stage1[d], m->st[d]));
    }
    for (int d = 0; d < n; ++d) {
        MCUDA_TRY(cudaSetDevice(d));
        for (int i = 0; i < n; ++i) {
            const int p = (d + i) % n;                        // own tails first, then the ring of peers
            if (p != d) MCUDA_TRY(cudaStreamWaitEvent(m->st[d], m->ev_stage1[p], 0));
            MCUDA_TRY(cudaMemcpyPeerAsync((char*)m->gathered[d] + (size_t)p * m->tail_bytes, d, m->tails[p], p, m->tail_bytes, m->st[d]));
        }
        MCUDA_TRY(cudaEventRecord(m->ev_pulled[d], m->st[d]));
        MRF_TRY(rf_plan_stage2(m->plan[d], m->in[d], m->out[d], m->gathered[d], n, d, m->st[d]));
    }
    return RF_OK;
}

static int mgpu_sync(rf_mgpu* m)
{
    for (int d = 0; d < m->n; ++d) { MCUDA_TRY(cudaSetDevice(d)); MCUDA_TRY(cudaStreamSynchronize(m->st[d])); }
    for (int d = 0; d < m->n; ++d) { MCUDA_TRY(cudaSetDevice(d)); MRF_TRY(rf_plan_check(m->plan[d])); }
    MCUDA_TRY(cudaSetDevice(m->home));
    return RF_OK;
}

int rf_mgpu_execute_host(rf_mgpu* m, const void* in_host, void* out_host)
{
    if (!m || !in_host || !out_host) return mfail(RF_EINVAL, "null argument");
    if (m->part_bytes == 0) return RF_OK;
    for (int d = 0; d < m->n; ++d) {
        MCUDA_TRY(cudaSetDevice(d));
        MCUDA_TRY(cudaMemcpyAsync(m->in[d], (const char*)in_host + (size_t)d * m->part_bytes, m->part_bytes, cudaMemcpyHostToDevice, m->st[d]));
    }
    int rc = mgpu_run(m);
    if (rc) { cudaSetDevice(m->home); return rc; }
    for (int d = 0; d < m->n; ++d) {
        MCUDA_TRY(cudaSetDevice(d));
        MCUDA_TRY(cudaMemcpyAsync((char*)out_host + (size_t)d * m->part_bytes, m->out[d], m->part_bytes, cudaMemcpyDeviceToHost, m->st[d]));
    }
    return mgpu_sync(m);
}

int rf_mgpu_profile(rf_mgpu* m, const void* in_host, int iters, float* ms_per_iter)
{
    if (!m || !in_host || !ms_per_iter || iters < 1) return mfail(RF_EINVAL, "bad argument");
    for (int d = 0; d < m->n; ++d) {
        MCUDA_TRY(cudaSetDevice(d));
        MCUDA_TRY(cudaMemcpyAsync(m->in[d], (const char*)in_host + (size_t)d * m->part_bytes, m->part_bytes, cudaMemcpyHostToDevice, m->st[d]));
    }
    int rc = mgpu_run(m);                                     // warm-up (lib/recfilter.cpp:995-997)
    if (rc == RF_OK) rc = mgpu_sync(m);
    if (rc) return rc;
    for (int d = 0; d < m->n; ++d) { MCUDA_TRY(cudaSetDevice(d)); MCUDA_TRY(cudaEventRecord(m->ev_t0[d], m->st[d])); }
    for (int i = 0; i < iters && rc == RF_OK; ++i) rc = mgpu_run(m);
    if (rc) { cudaSetDevice(m->home); return rc; }
    for (int d = 0; d < m->n; ++d) { MCUDA_TRY(cudaSetDevice(d)); MCUDA_TRY(cudaEventRecord(m->ev_t1[d], m->st[d])); }
    if ((rc = mgpu_sync(m))) return rc;
    float worst = 0.f;
    for (int d = 0; d < m->n; ++d) {
        float t = 0.f;
        MCUDA_TRY(cudaSetDevice(d));
        MCUDA_TRY(cudaEventElapsedTime(&t, m->ev_t0[d], m->ev_t1[d]));
        if (t > worst) worst = t;
    }
    MCUDA_TRY(cudaSetDevice(m->home));
    *ms_per_iter = worst / iters;                             // device time, the slowest GPU
    return RF_OK;
}

#pragma GCC visibility pop
} // extern "C"
