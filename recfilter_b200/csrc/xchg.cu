/*
 * xchg.cu -- exchange windows for the strip tails of a sharded filter: peer-to-peer over NVLink, no
 * collective library and no host in the data path (C ABI: rf_xchg_*, include/recfilter_b200.h).
 *
 * The reference is single-GPU; this is the exchange step of SURVEY 8e ("one ncclAllGather (or P2P put +
 * flag) of those [tails]").  Every rank owns a WINDOW in its device memory: two generations (double buffer)
 * of `nranks` slots of `bytes_per_rank`, plus one arrival word per slot.  A step is
 *
 *   put    my tails -> slot[my rank] of EVERY rank's window (cudaMemcpyAsync through the peer mapping, in
 *          stream order), then a one-warp kernel stores the step number into the arrival word of my slot in
 *          every window (a peer store after the copies: the payload is in place when the word changes);
 *   wait   a one-warp kernel on the consumer's stream spins until all arrival words of the generation show
 *          the step number (it gives up after a bounded number of polls and raises the error word instead
 *          of hanging the device); the window generation then is the all-gathered [nranks][bytes_per_rank]
 *          array rf_plan_stage2 takes.
 *
 * Two generations suffice: a rank can only start step k+2 after every rank has put step k+1, which every
 * rank does after it has consumed generation k (its own stream order).
 * Windows of other processes (one process per GPU, torchrun) are mapped through CUDA IPC handles that the
 * host layer passes around once at set-up; in a single process the peer window pointer is used directly.
 */
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "../../include/recfilter_b200.h"

namespace {

thread_local char x_err[512] = "";
int xfail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(x_err, sizeof(x_err), fmt, ap);
    va_end(ap);
    return code;
}
#define XCUDA_TRY(expr)                                                                  \
    do { cudaError_t e__ = (expr);                                                       \
         if (e__ != cudaSuccess)                                                         \
             return xfail(RF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

constexpr int XCHG_MAX_RANKS = 32;
constexpr unsigned XCHG_SPIN_LIMIT = 1u << 24;

struct PeerWords { unsigned* w[XCHG_MAX_RANKS]; };

// thread p: my arrival word in rank p's window <- step (system scope: the word lives in another GPU's memory)
__global__ void xchg_signal_kernel(PeerWords words, int nranks, unsigned step)
{
    const int p = threadIdx.x;
    if (p < nranks && words.w[p]) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(words.w[p]), "r"(step) : "memory");
    }
}
// thread r: wait for rank r's slot of this generation
__global__ void xchg_wait_kernel(const unsigned* words, int nranks, unsigned step, unsigned* err, unsigned mask)
{
    const int r = threadIdx.x;
    if (r >= nranks || !((mask >> r) & 1u)) return;
    unsigned spins = 0;
    while (true) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(words + r) : "memory");
        if (v == step) break;
        if (++spins > XCHG_SPIN_LIMIT) { *err = 1u; break; }
        __nanosleep(100);
    }
}

} // namespace

struct rf_xchg {
    int nranks = 1, rank = 0, device = 0;
    size_t bytes = 0;                 // per rank and generation
    size_t gen_bytes = 0;             // nranks * bytes rounded up to 256
    size_t words_off = 0;             // arrival words: [2][XCHG_MAX_RANKS]
    size_t total = 0;
    unsigned char* win = nullptr;     // my window
    unsigned char* peer[XCHG_MAX_RANKS] = {};   // every rank's window as seen from this process (peer[rank] == win)
    bool ipc_opened[XCHG_MAX_RANKS] = {};
    unsigned* err = nullptr;          // device error word
    unsigned step = 0;                // last step put
    unsigned waited = 0;              // last step waited for
    bool open = false;                // a step is being assembled from parts (rf_xchg_put_part)
    unsigned targets = 0;             // ranks whose windows received a part of the open step
};

extern "C" {
#pragma GCC visibility push(default)

const char* rf_xchg_last_error(void) { return x_err; }

int rf_xchg_create(size_t bytes_per_rank, int nranks, int rank, rf_xchg** out)
{
    if (!out) return xfail(RF_EINVAL, "null argument");
    *out = nullptr;
    if (nranks < 1 || nranks > XCHG_MAX_RANKS || rank < 0 || rank >= nranks) return xfail(RF_EINVAL, "bad rank / nranks (<= %d)", XCHG_MAX_RANKS);
    if (bytes_per_rank == 0 || (bytes_per_rank & 3)) return xfail(RF_EINVAL, "bytes_per_rank must be a positive multiple of 4");
    rf_xchg* x = new (std::nothrow) rf_xchg();
    if (!x) return xfail(RF_ENOMEM, "out of host memory");
    x->nranks = nranks; x->rank = rank; x->bytes = bytes_per_rank;
    XCUDA_TRY(cudaGetDevice(&x->device));
    x->gen_bytes = ((size_t)nranks * bytes_per_rank + 255) / 256 * 256;
    x->words_off = 2 * x->gen_bytes;
    x->total = x->words_off + 2 * XCHG_MAX_RANKS * sizeof(unsigned) + 256;
    XCUDA_TRY(cudaMalloc((void**)&x->win, x->total));
    XCUDA_TRY(cudaMemset(x->win, 0, x->total));
    XCUDA_TRY(cudaDeviceSynchronize());
    x->err = reinterpret_cast<unsigned*>(x->win + x->words_off + 2 * XCHG_MAX_RANKS * sizeof(unsigned));
    x->peer[rank] = x->win;
    *out = x;
    return RF_OK;
}

void rf_xchg_destroy(rf_xchg* x)
{
    if (!x) return;
    for (int p = 0; p < x->nranks; ++p)
        if (x->ipc_opened[p] && x->peer[p]) cudaIpcCloseMemHandle(x->peer[p]);
    if (x->win) cudaFree(x->win);
    delete x;
}

size_t rf_xchg_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int rf_xchg_ipc_handle(rf_xchg* x, void* handle)
{
    if (!x || !handle) return xfail(RF_EINVAL, "null argument");
    cudaIpcMemHandle_t h;
    XCUDA_TRY(cudaIpcGetMemHandle(&h, x->win));
    std::memcpy(handle, &h, sizeof(h));
    return RF_OK;
}

int rf_xchg_open_peer(rf_xchg* x, int peer, const void* handle)
{
    if (!x || !handle || peer < 0 || peer >= x->nranks) return xfail(RF_EINVAL, "bad argument");
    if (peer == x->rank) return RF_OK;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    XCUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer] = (unsigned char*)p;
    x->ipc_opened[peer] = true;
    return RF_OK;
}

int rf_xchg_set_peer(rf_xchg* x, int peer, rf_xchg* other)
{
    if (!x || !other || peer < 0 || peer >= x->nranks || other->rank != peer) return xfail(RF_EINVAL, "bad argument");
    if (peer == x->rank) return RF_OK;
    if (other->device != x->device) {
        int can = 0;
        XCUDA_TRY(cudaDeviceCanAccessPeer(&can, x->device, other->device));
        if (!can) return xfail(RF_EUNSUPPORTED, "device %d cannot access device %d peer to peer", x->device, other->device);
        int cur = 0;
        XCUDA_TRY(cudaGetDevice(&cur));
        XCUDA_TRY(cudaSetDevice(x->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        cudaSetDevice(cur);
        XCUDA_TRY(e);
    }
    x->peer[peer] = other->win;
    return RF_OK;
}

int rf_xchg_put(rf_xchg* x, const void* src_dev, size_t bytes, void* stream)
{
    if (!x || !src_dev) return xfail(RF_EINVAL, "null argument");
    if (bytes > x->bytes) return xfail(RF_EINVAL, "put of %zu bytes into slots of %zu", bytes, x->bytes);
    if (x->open) return xfail(RF_EINVAL, "rf_xchg_put inside a step opened by rf_xchg_put_part");
    for (int p = 0; p < x->nranks; ++p)
        if (!x->peer[p]) return xfail(RF_EINVAL, "window of rank %d has not been mapped", p);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned step = ++x->step;
    const size_t gen = step & 1u;
    PeerWords words;
    std::memset(&words, 0, sizeof(words));
    // peers first, the own window last (ring order: every rank starts with a different peer)
    for (int i = 1; i <= x->nranks; ++i) {
        const int p = (x->rank + i) % x->nranks;
        XCUDA_TRY(cudaMemcpyAsync(x->peer[p] + gen * x->gen_bytes + (size_t)x->rank * x->bytes, src_dev, bytes, cudaMemcpyDefault, st));
        words.w[p] = reinterpret_cast<unsigned*>(x->peer[p] + x->words_off) + gen * XCHG_MAX_RANKS + x->rank;
    }
    xchg_signal_kernel<<<1, 32, 0, st>>>(words, x->nranks, step);
    XCUDA_TRY(cudaGetLastError());
    return RF_OK;
}

int rf_xchg_wait(rf_xchg* x, void* stream, void** gathered_dev)
{
    if (!x) return xfail(RF_EINVAL, "null argument");
    if (x->step == x->waited) return xfail(RF_EINVAL, "rf_xchg_wait without a put of this step");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned step = x->step;
    const size_t gen = step & 1u;
    xchg_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const unsigned*>(x->win + x->words_off) + gen * XCHG_MAX_RANKS, x->nranks, step, x->err, 0xffffffffu);
    XCUDA_TRY(cudaGetLastError());
    x->waited = step;
    if (gathered_dev) *gathered_dev = x->win + gen * x->gen_bytes;
    return RF_OK;
}

// One step assembled from parts: bytes [offset, offset + bytes) of my tails go into my slot of the windows of the ranks in
// `peer_mask` (bit p = rank p; my own bit = my own window); the call with last != 0 raises my arrival word in every
// window that received a part.  Parts of a slot that are never written stay zero.
int rf_xchg_put_part(rf_xchg* x, unsigned peer_mask, const void* src_dev, size_t offset, size_t bytes, int last, void* stream)
{
    if (!x) return xfail(RF_EINVAL, "null argument");
    if (offset + bytes > x->bytes || (bytes && !src_dev)) return xfail(RF_EINVAL, "part [%zu, %zu) outside slots of %zu bytes", offset, offset + bytes, x->bytes);
    if (x->nranks < 32) peer_mask &= (1u << x->nranks) - 1u;
    for (int p = 0; p < x->nranks; ++p)
        if (((peer_mask >> p) & 1u) && !x->peer[p]) return xfail(RF_EINVAL, "window of rank %d has not been mapped", p);
    cudaStream_t st = (cudaStream_t)stream;
    if (!x->open) { ++x->step; x->open = true; x->targets = 0; }
    const unsigned step = x->step;
    const size_t gen = step & 1u;
    if (bytes)
        for (int i = 1; i <= x->nranks; ++i) {
            const int p = (x->rank + i) % x->nranks;
            if (!((peer_mask >> p) & 1u)) continue;
            XCUDA_TRY(cudaMemcpyAsync(x->peer[p] + gen * x->gen_bytes + (size_t)x->rank * x->bytes + offset,
                                      (const unsigned char*)src_dev + offset, bytes, cudaMemcpyDefault, st));
        }
    x->targets |= peer_mask;
    if (last) {
        PeerWords words;
        std::memset(&words, 0, sizeof(words));
        for (int p = 0; p < x->nranks; ++p)
            if ((x->targets >> p) & 1u) words.w[p] = reinterpret_cast<unsigned*>(x->peer[p] + x->words_off) + gen * XCHG_MAX_RANKS + x->rank;
        xchg_signal_kernel<<<1, 32, 0, st>>>(words, x->nranks, step);
        XCUDA_TRY(cudaGetLastError());
        x->open = false;
    }
    return RF_OK;
}

// as rf_xchg_wait, for the slots of the ranks in `from_mask` only
int rf_xchg_wait_from(rf_xchg* x, unsigned from_mask, void* stream, void** gathered_dev)
{
    if (!x) return xfail(RF_EINVAL, "null argument");
    if (x->open) return xfail(RF_EINVAL, "rf_xchg_wait_from inside an open step (the last part was not marked)");
    if (x->step == x->waited) return xfail(RF_EINVAL, "rf_xchg_wait_from without a put of this step");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned step = x->step;
    const size_t gen = step & 1u;
    xchg_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const unsigned*>(x->win + x->words_off) + gen * XCHG_MAX_RANKS, x->nranks, step, x->err, from_mask);
    XCUDA_TRY(cudaGetLastError());
    x->waited = step;
    if (gathered_dev) *gathered_dev = x->win + gen * x->gen_bytes;
    return RF_OK;
}

int rf_xchg_check(rf_xchg* x)
{
    if (!x) return xfail(RF_EINVAL, "null argument");
    unsigned v = 0;
    XCUDA_TRY(cudaMemcpy(&v, x->err, sizeof(v), cudaMemcpyDeviceToHost));
    if (v) return xfail(RF_EINTERNAL, "a rank's tails never arrived (exchange wait gave up)");
    return RF_OK;
}

#pragma GCC visibility pop
} // extern "C"
