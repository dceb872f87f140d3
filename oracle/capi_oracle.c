/*
 * capi_oracle.c -- the C ABI of include/recfilter_b200.h implemented on the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists for one purpose: to link the reference's own test
 * programs (/root/reference/tests/*.cpp, compiled unchanged against include/recfilter.h) with
 * the oracle instead of the CUDA engine, so that the inline serial loops those programs carry
 * pin the oracle (oracle/pin_reference.py -> tests/golden/PIN_REPORT.json).  The product
 * (recfilter_b200/) never links or loads this file; "device" pointers here are host pointers.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/recfilter_b200.h"

int oracle_filter(void* data, int dtype, int ndim, const int64_t* extents, int clamp, int nscans,
                  const int* scan_dim, const int* scan_causal, const int* scan_ncoeff, const float* coeffs,
                  int nthreads);

struct rf_plan { rf_desc d; size_t total; size_t eb; };

static char g_err[256] = "";
static size_t elem_bytes(int dt)
{
    switch (dt) { case RF_F32: case RF_I32: case RF_U32: return 4; case RF_I16: case RF_U16: return 2;
                  case RF_I8: case RF_U8: return 1; }
    return 0;
}

const char* rf_version(void) { return "recfilter oracle backend (CPU, test infrastructure)"; }
const char* rf_last_error(void) { return g_err; }
int rf_device_count(void) { return 1; }
int rf_set_device(int device) { (void)device; return RF_OK; }

void oracle_set_sum_order(int order);

int rf_plan_create(const rf_desc* desc, rf_plan** out)
{
    const char* so = getenv("ORACLE_SUM_ORDER");                 /* "tests": see oracle.c */
    oracle_set_sum_order(so && strcmp(so, "tests") == 0);
    if (!desc || !out) { snprintf(g_err, sizeof g_err, "null argument"); return RF_EINVAL; }
    /* the checker keeps the pointwise stage separate (the host then runs it through rf_stencil_execute) */
    if (desc->opt.epilogue) { snprintf(g_err, sizeof g_err, "oracle backend: no fused epilogue"); return RF_EUNSUPPORTED; }
    rf_plan* p = (rf_plan*)calloc(1, sizeof(rf_plan));
    if (!p) return RF_ENOMEM;
    p->d = *desc;
    p->eb = elem_bytes(desc->dtype);
    if (!p->eb) { free(p); snprintf(g_err, sizeof g_err, "unknown dtype"); return RF_EINVAL; }
    p->total = 1;
    for (int i = 0; i < desc->ndim; ++i) p->total *= (size_t)desc->extent[i];
    for (int s = 0; s < desc->nscans; ++s)
        if (desc->scans[s].dim < 0 || desc->scans[s].dim >= desc->ndim || desc->scans[s].order < 1) {
            free(p); snprintf(g_err, sizeof g_err, "bad scan %d", s); return RF_EINVAL;
        }
    *out = p;
    return RF_OK;
}
void rf_plan_destroy(rf_plan* plan) { free(plan); }
size_t rf_plan_workspace_bytes(const rf_plan* plan) { (void)plan; return 0; }
int rf_plan_num_launches(const rf_plan* plan) { (void)plan; return 0; }
int rf_plan_describe(const rf_plan* plan, char* buf, size_t n)
{
    snprintf(buf, n, "oracle backend: %d scans applied serially on the CPU (test infrastructure)\n", plan->d.nscans);
    return RF_OK;
}

/* oracle.c numbers the element types like rf_dtype (F32 0, I32 2, U32 3, I16 4, U16 5, I8 6, U8 7) */
int rf_plan_execute(rf_plan* plan, const void* in_dev, void* out_dev, void* stream)
{
    (void)stream;
    const rf_desc* d = &plan->d;
    if (plan->total == 0) return RF_OK;
    if (in_dev != out_dev) memmove(out_dev, in_dev, plan->total * plan->eb);
    int dims[RF_MAX_SCANS], causal[RF_MAX_SCANS], nco[RF_MAX_SCANS];
    float* co = (float*)malloc(sizeof(float) * RF_MAX_SCANS * (RF_MAX_ORDER + 1));
    size_t k = 0;
    for (int s = 0; s < d->nscans; ++s) {
        dims[s] = d->scans[s].dim; causal[s] = d->scans[s].causal; nco[s] = d->scans[s].order + 1;
        for (int j = 0; j <= d->scans[s].order; ++j) co[k++] = d->scans[s].coeff[j];
    }
    int rc = oracle_filter(out_dev, d->dtype, d->ndim, d->extent, d->border == RF_BORDER_CLAMP, d->nscans,
                           dims, causal, nco, co, 1);
    free(co);
    if (rc) { snprintf(g_err, sizeof g_err, "oracle_filter failed"); return RF_EINVAL; }
    return RF_OK;
}
int rf_plan_execute_host(rf_plan* plan, const void* in_host, void* out_host) { return rf_plan_execute(plan, in_host, out_host, 0); }
/* pointwise linear stencil (plain loops; the reference's box_filter.h differencing Func) */
int rf_stencil_execute(int ndim, const int64_t* extent, int dtype, int ntaps, const rf_tap* taps, float post_scale,
                       const void* in_dev, const void* in2_dev, void* out_dev, void* stream)
{
    (void)stream;
    if (ndim < 1 || ndim > RF_MAX_DIMS || ntaps < 1 || ntaps > RF_MAX_TAPS || !in_dev || !out_dev || in_dev == out_dev) return RF_EINVAL;
    if (dtype != RF_F32 && dtype != RF_I32 && dtype != RF_U32) return RF_EUNSUPPORTED;
    int64_t ext[RF_MAX_DIMS] = { 1, 1, 1, 1 }, total = 1;
    for (int d = 0; d < ndim; ++d) { ext[d] = extent[d]; total *= extent[d]; }
    for (int64_t i = 0; i < total; ++i) {
        int64_t c[RF_MAX_DIMS], r = i;
        for (int d = 0; d < RF_MAX_DIMS; ++d) { c[d] = r % ext[d]; r /= ext[d]; }
        float accf = 0.0f; uint32_t accu = 0u;
        for (int t = 0; t < ntaps; ++t) {
            int64_t idx = 0, stride = 1;
            for (int d = 0; d < ndim; ++d) {
                int64_t v = c[d] + taps[t].offset[d];
                if (v > taps[t].hi[d]) v = taps[t].hi[d];
                if (v < taps[t].lo[d]) v = taps[t].lo[d];
                if (v > ext[d] - 1) v = ext[d] - 1;
                if (v < 0) v = 0;
                idx += v * stride; stride *= ext[d];
            }
            const void* src = taps[t].source ? in2_dev : in_dev;
            if (!src) return RF_EINVAL;
            if (dtype == RF_F32) accf = accf + taps[t].weight * ((const float*)src)[idx];
            else accu = accu + (uint32_t)(int32_t)(taps[t].weight < 0 ? taps[t].weight - 0.5f : taps[t].weight + 0.5f) * ((const uint32_t*)src)[idx];
        }
        if (dtype == RF_F32) ((float*)out_dev)[i] = accf * post_scale;
        else ((uint32_t*)out_dev)[i] = accu * (uint32_t)(int32_t)(post_scale < 0 ? post_scale - 0.5f : post_scale + 0.5f);
    }
    return RF_OK;
}

int rf_plan_execute_host_batch(rf_plan* plan, int n, const void* const* in_host, void* const* out_host)
{
    for (int i = 0; i < n; ++i) { int rc = rf_plan_execute(plan, in_host[i], out_host[i], 0); if (rc) return rc; }
    return RF_OK;
}

int rf_plan_profile(rf_plan* plan, const void* in_dev, void* out_dev, int iters, float* ms)
{
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int i = 0; i < iters; ++i) { int rc = rf_plan_execute(plan, in_dev, out_dev, 0); if (rc) return rc; }
    clock_gettime(CLOCK_MONOTONIC, &b);
    *ms = (float)(((b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6) / iters);
    return RF_OK;
}
int rf_plan_stage_timing(rf_plan* plan, int enable) { (void)plan; (void)enable; return RF_OK; }
int rf_plan_stage_times(rf_plan* plan, double* ms, long* counts, int n)
{
    (void)plan;
    for (int i = 0; i < n; ++i) { ms[i] = 0; counts[i] = 0; }
    return RF_OK;
}
size_t rf_plan_shard_tail_bytes(const rf_plan* plan) { (void)plan; return 0; }
int rf_plan_stage1(rf_plan* p, const void* i, void* o, void* t, void* s) { (void)p; (void)i; (void)o; (void)t; (void)s; return RF_EUNSUPPORTED; }
int rf_plan_stage2(rf_plan* p, const void* i, void* o, const void* g, int n, int r, void* s)
{ (void)p; (void)i; (void)o; (void)g; (void)n; (void)r; (void)s; return RF_EUNSUPPORTED; }

int rf_plan_shard_vectors(const rf_plan* p) { (void)p; return 0; }
int rf_plan_shard_neighbors_suffice(const rf_plan* p) { (void)p; return 0; }
int rf_plan_shard_resolve_lines(rf_plan* p, const void* g, int n, int64_t l, void* e, void* s)
{ (void)p; (void)g; (void)n; (void)l; (void)e; (void)s; return RF_EUNSUPPORTED; }
int rf_plan_stage2_ext(rf_plan* p, const void* i, void* o, const void* e, void* s)
{ (void)p; (void)i; (void)o; (void)e; (void)s; return RF_EUNSUPPORTED; }
int rf_clock_begin(void* stream, void** clock)
{
    (void)stream;
    struct timespec* t = (struct timespec*)malloc(sizeof(struct timespec));
    clock_gettime(CLOCK_MONOTONIC, t);
    *clock = t;
    return RF_OK;
}
int rf_clock_end(void* clock, void* stream, float* ms)
{
    (void)stream;
    struct timespec b, *a = (struct timespec*)clock;
    clock_gettime(CLOCK_MONOTONIC, &b);
    *ms = (float)((b.tv_sec - a->tv_sec) * 1e3 + (b.tv_nsec - a->tv_nsec) * 1e-6);
    free(a);
    return RF_OK;
}

int rf_malloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? RF_OK : RF_ENOMEM; }
int rf_free(void* p) { free(p); return RF_OK; }
int rf_memcpy_h2d(void* d, const void* s, size_t n) { memcpy(d, s, n); return RF_OK; }
int rf_memcpy_d2h(void* d, const void* s, size_t n) { memcpy(d, s, n); return RF_OK; }
int rf_memset(void* d, int v, size_t n) { memset(d, v, n); return RF_OK; }
int rf_malloc_host(void** p, size_t bytes) { return rf_malloc(p, bytes); }
int rf_free_host(void* p) { free(p); return RF_OK; }
int rf_synchronize(void) { return RF_OK; }

/* entry points added in round 2: the oracle backend is one CPU "device" -- no look-back kernels to check, no peers */
int rf_plan_check(rf_plan* p) { (void)p; return RF_OK; }
int rf_mgpu_create(const rf_desc* d, int n, rf_mgpu** out)
{
    (void)d; (void)n;
    if (out) *out = 0;
    snprintf(g_err, sizeof(g_err), "the oracle backend is a single CPU device");
    return RF_EUNSUPPORTED;
}
void rf_mgpu_destroy(rf_mgpu* m) { (void)m; }
int rf_mgpu_ngpus(const rf_mgpu* m) { (void)m; return 1; }
int rf_mgpu_describe(const rf_mgpu* m, char* buf, size_t n) { (void)m; if (buf && n) buf[0] = 0; return RF_EUNSUPPORTED; }
int rf_mgpu_execute_host(rf_mgpu* m, const void* i, void* o) { (void)m; (void)i; (void)o; return RF_EUNSUPPORTED; }
int rf_mgpu_profile(rf_mgpu* m, const void* i, int it, float* ms) { (void)m; (void)i; (void)it; (void)ms; return RF_EUNSUPPORTED; }
const char* rf_mgpu_last_error(void) { return g_err; }
size_t rf_xchg_handle_bytes(void) { return 64; }
int rf_xchg_create(size_t b, int n, int r, rf_xchg** out) { (void)b; (void)n; (void)r; if (out) *out = 0; return RF_EUNSUPPORTED; }
void rf_xchg_destroy(rf_xchg* x) { (void)x; }
int rf_xchg_ipc_handle(rf_xchg* x, void* h) { (void)x; (void)h; return RF_EUNSUPPORTED; }
int rf_xchg_open_peer(rf_xchg* x, int p, const void* h) { (void)x; (void)p; (void)h; return RF_EUNSUPPORTED; }
int rf_xchg_set_peer(rf_xchg* x, int p, rf_xchg* o) { (void)x; (void)p; (void)o; return RF_EUNSUPPORTED; }
int rf_xchg_put(rf_xchg* x, const void* s, size_t b, void* st) { (void)x; (void)s; (void)b; (void)st; return RF_EUNSUPPORTED; }
int rf_xchg_wait(rf_xchg* x, void* st, void** g) { (void)x; (void)st; (void)g; return RF_EUNSUPPORTED; }
int rf_xchg_put_part(rf_xchg* x, unsigned m, const void* s, size_t o, size_t b, int l, void* st) { (void)x; (void)m; (void)s; (void)o; (void)b; (void)l; (void)st; return RF_EUNSUPPORTED; }
int rf_xchg_wait_from(rf_xchg* x, unsigned m, void* st, void** g) { (void)x; (void)m; (void)st; (void)g; return RF_EUNSUPPORTED; }
int rf_xchg_check(rf_xchg* x) { (void)x; return RF_EUNSUPPORTED; }
const char* rf_xchg_last_error(void) { return g_err; }
