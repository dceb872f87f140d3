/*
 * recfilter_b200.h -- C ABI of the B200-native recursive-filter engine.
 *
 * This is the drop-in boundary: the C++ operator surface in include/recfilter.h
 * (RecFilter, add_filter, split, realize, profile ...) is implemented on top of
 * exactly these entry points, and a maintainer of mit-gfx/recfilter would bind
 * the same functions from lib/recfilter.cpp.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference):
 *
 *   rf_plan_create      RecFilter::define + add_filter + split + finalize/compile_jit
 *                       lib/recfilter.cpp:192-248, 260-392, 918-930; lib/split.cpp:1850-2080
 *                       (scan list + tiling -> executable pipeline; here a launch plan
 *                       with carry matrices instead of a Halide Func DAG)
 *   rf_plan_execute     Func::realize on device buffers, lib/recfilter.cpp:984-989, 998-1009
 *   rf_plan_execute_host  create_realization's copy_to_dev + realize + Image<T>(Realization)
 *                       download, lib/recfilter.cpp:932-982, halide/src/runtime/cuda.cpp:512,575
 *   rf_plan_profile     RecFilter::profile, lib/recfilter.cpp:991-1016 (CUDA events instead of
 *                       an unsynchronised wall clock)
 *   rf_plan_describe    operator<<(RecFilter) / print_synopsis, lib/recfilter.cpp:1024-1096
 *   rf_plan_stage1/2, rf_plan_shard_*   no reference equivalent (the reference is single-GPU);
 *                       strip-sharded execution with an order-r carry exchange
 *
 * Conventions: plain C types only; every function returns 0 on success and a
 * negative RF_E* code on failure, with a message retrievable by rf_last_error();
 * no exceptions cross the boundary.  A plan is immutable after creation; execute
 * may be called repeatedly.  Arrays are dense with dimension 0 contiguous (the
 * Halide::Image layout, lib/recfilter.cpp:970-981).
 */
#ifndef RECFILTER_B200_H_
#define RECFILTER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_MAX_DIMS   4
#define RF_MAX_SCANS  64
#define RF_MAX_ORDER  32

/* element types (type of the filter = type of the RHS, lib/recfilter.cpp:197) */
enum rf_dtype {
    RF_F32 = 0,
    RF_I32 = 2,
    RF_U32 = 3,
    RF_I16 = 4,
    RF_U16 = 5,
    RF_I8  = 6,
    RF_U8  = 7
};

/* image border (RecFilter::set_clamped_image_border, lib/recfilter.cpp:252-258) */
enum rf_border { RF_BORDER_ZERO = 0, RF_BORDER_CLAMP = 1 };

/* tile engines: the fused fast path (128x128 / 64x64 register tiles, orders <= 4 -- 8 for single-scan long
 * signals --, partial last tiles when the row pitch is a multiple of 16 bytes) and the generic engine (any
 * order <= 32, any extents, honoured split() tiles) */
enum rf_engine { RF_ENGINE_AUTO = 0, RF_ENGINE_GENERIC = 1, RF_ENGINE_FUSED = 2,
                 RF_ENGINE_TWOPASS = 3   /* fused tile engine, but never the single-pass look-back kernels */ };

/* error codes */
enum rf_status {
    RF_OK            =  0,
    RF_EINVAL        = -1,   /* bad descriptor / argument */
    RF_EUNSUPPORTED  = -2,   /* valid request this engine does not implement */
    RF_ECUDA         = -3,   /* CUDA runtime error (message has the CUDA string) */
    RF_ENOMEM        = -4,
    RF_ENODEVICE     = -5,   /* no CUDA device: there is no CPU fallback */
    RF_EINTERNAL     = -6    /* a kernel reported an internal failure (rf_plan_check) */
};

/* one scan = one RecFilter::add_filter call (lib/recfilter.cpp:264-392) */
typedef struct rf_scan {
    int32_t dim;                      /* dimension scanned, 0 = contiguous */
    int32_t causal;                   /* 1: +x (i ascending), 0: -x */
    int32_t order;                    /* r = number of feedback coefficients, 1..RF_MAX_ORDER */
    float   coeff[RF_MAX_ORDER + 1];  /* {b0, a1..ar}: feedback terms are ADDED */
} rf_scan;

/* planner options; zero-initialised means "engine decides" */
typedef struct rf_options {
    int32_t tile[RF_MAX_DIMS];  /* split() hint per dimension; 0 = engine picks (64).
                                   Values are clamped to the kernel's register tile. */
    int32_t honor_tile;         /* 1: use tile[] literally (<= register tile) even if small */
    int32_t fuse_dims;          /* 1: fuse scans of dimension 0 and 1 in one pass (default),
                                   0: one pass per dimension (cascade_by_dimension style),
                                   -1: engine decides */
    int32_t open_lo;            /* sharding: the low / high face of shard_dim is an interior */
    int32_t open_hi;            /*   cut, carries come from the neighbour shard */
    int32_t shard_dim;          /* dimension that is sharded across devices, -1: none */
    int32_t engine;             /* rf_engine: which tile engine a pass may use (0: planner decides) */
    int32_t epilogue;           /* 1: out = epi_out * filtered + epi_in * input, fused into the filter's last store -- the
                                   pointwise stage the reference merges with RecFilter::compute_at (unsharp mask,
                                   apps/usm/unsharp_mask_optimized.cpp:61-66).  RF_EUNSUPPORTED unless the filter is one
                                   fused float pass with scans along dimension 0; in-place execution stays legal */
    float   epi_in, epi_out;
    int32_t reserved[4];
} rf_options;

/* filter descriptor */
typedef struct rf_desc {
    int32_t ndim;                       /* 1..RF_MAX_DIMS */
    int64_t extent[RF_MAX_DIMS];        /* samples per dimension, extent[0] contiguous */
    int32_t dtype;                      /* rf_dtype */
    int32_t border;                     /* rf_border */
    int32_t nscans;                     /* 0..RF_MAX_SCANS, applied in this order */
    rf_scan scans[RF_MAX_SCANS];
    rf_options opt;
} rf_desc;

typedef struct rf_plan rf_plan;         /* opaque */

/* library / device */
const char* rf_version(void);
const char* rf_last_error(void);
int  rf_device_count(void);
int  rf_set_device(int device);

/* plan life cycle */
int    rf_plan_create(const rf_desc* desc, rf_plan** out);
void   rf_plan_destroy(rf_plan* plan);
size_t rf_plan_workspace_bytes(const rf_plan* plan);
int    rf_plan_num_launches(const rf_plan* plan);           /* kernels per execute */
int    rf_plan_describe(const rf_plan* plan, char* buf, size_t n);

/*
 * Run the filter.  in_dev/out_dev are device pointers to dense arrays of the
 * plan's dtype; they may alias (in-place).  stream is a cudaStream_t passed as
 * void* (NULL = default stream).  Asynchronous with respect to the host.
 */
int rf_plan_execute(rf_plan* plan, const void* in_dev, void* out_dev, void* stream);

/*
 * Wait for the device and report whether any kernel of the plan flagged an internal failure since the plan was
 * created (the single-pass look-back kernels give up, instead of hanging, when a predecessor tile never
 * publishes its carry).  RF_OK, or RF_EINTERNAL.  The host-buffer entry points and rf_plan_profile call it.
 */
int rf_plan_check(rf_plan* plan);

/* Host buffers: H2D + execute + D2H, synchronous (the realize() path). */
int rf_plan_execute_host(rf_plan* plan, const void* in_host, void* out_host);

/*
 * n independent images/signals of the plan's shape in host memory (the app loops that call
 * realize() once per frame, e.g. /root/reference/lib/recfilter.cpp:991-1016 with a fresh input):
 * upload, filter and download are pipelined over three device buffers, so the H2D copy of image
 * i+1 and the D2H copy of image i-1 overlap the kernels of image i.  Pinned host memory
 * (rf_malloc_host) is needed for the copies to overlap; pageable memory works but serialises.
 * Synchronous: all outputs are complete on return.
 */
int rf_plan_execute_host_batch(rf_plan* plan, int n, const void* const* in_host, void* const* out_host);

/*
 * Pointwise epilogue: out(x) = post_scale * sum_i weight_i * in(clamp(x + offset_i, lo_i, hi_i)) per dimension
 * (the common factor is applied after the sum, as "(a - b - c + d) / area" does: the taps of a summed-area table
 * are large numbers whose differences must be taken before scaling) -- the finite
 * differencing that turns a summed-area table into a box filter (apps/box/box_filter.h:36-39, 128-139 in the
 * reference, where it is a separate Halide Func scheduled with compute_root) and similar small linear stencils
 * of one filter result, or of two arrays of the same shape (unsharp mask: (1+w)*image - w*blur,
 * apps/usm/unsharp_mask_naive.cpp:61; in2_dev may be NULL when no tap reads source 1).  Indices are
 * additionally clamped to the array.  dtype: RF_F32, RF_I32 or RF_U32
 * (integer weights are the rounded float weights, arithmetic wraps).  in_dev and out_dev must not alias.
 */
#define RF_MAX_TAPS 32
typedef struct rf_tap {
    float   weight;
    int32_t source;                     /* 0: in_dev, 1: in2_dev (e.g. the original image beside its blur) */
    int32_t offset[RF_MAX_DIMS];
    int32_t lo[RF_MAX_DIMS];            /* INT32_MIN: no lower clamp */
    int32_t hi[RF_MAX_DIMS];            /* INT32_MAX: no upper clamp */
} rf_tap;
int rf_stencil_execute(int ndim, const int64_t* extent, int dtype, int ntaps, const rf_tap* taps, float post_scale,
                       const void* in_dev, const void* in2_dev, void* out_dev, void* stream);

/* Time `iters` executions on device-resident data with CUDA events (ms per iteration). */
int rf_plan_profile(rf_plan* plan, const void* in_dev, void* out_dev, int iters, float* ms_per_iter);

/*
 * Per-stage device timing (development / bench aid; the reference's analogue is the
 * nvprof kernel-sum of scripts/cuda_profile.sh:27-35).  While enabled, every kernel of
 * rf_plan_execute / stage1 / stage2 is bracketed by CUDA events on its stream.
 * Stages: 0 tile kernel (tails), 1 carry chains, 2 cross-dimension residual,
 * 3 tile kernel (final), 4 integer widen/narrow.  ms[] are accumulated milliseconds and
 * counts[] the number of launches since timing was enabled.
 */
#define RF_NUM_STAGES 5
int rf_plan_stage_timing(rf_plan* plan, int enable);
int rf_plan_stage_times(rf_plan* plan, double* ms, long* counts, int n);

/*
 * Strip-sharded execution (one plan per device, opt.shard_dim set).
 *   stage1: intra-shard work with zero incoming carries (passes before the sharded
 *           one run to completion into out_dev); writes this shard's outgoing
 *           boundary tails to tails_dev (rf_plan_shard_tail_bytes bytes).
 *   the caller exchanges tails between devices (all-gather, rank order)
 *   stage2: consumes the gathered tails of all `nshards` shards
 *           (nshards * tail_bytes, rank-major), resolves this shard's incoming
 *           carries and finishes the filter.
 */
size_t rf_plan_shard_tail_bytes(const rf_plan* plan);
/*
 * Column-chunked exchange (less traffic than the all-gather when there are many shards): the tails are
 * [vectors][lines] with vectors = scans x order (rf_plan_shard_vectors) and lines = tail elements / vectors.
 * Every rank receives the tails of ALL shards for ITS chunk of the lines (all-to-all), resolves the carries
 * entering every shard for those lines (rf_plan_shard_resolve_lines: gathered is [nshards][vectors][nlines],
 * ext_all the same shape), returns each shard its chunk (all-to-all) and finishes with rf_plan_stage2_ext,
 * whose ext_dev is the [vectors][lines] array of carries entering this shard.
 */
int rf_plan_shard_vectors(const rf_plan* plan);
int rf_plan_shard_neighbors_suffice(const rf_plan* plan);   /* 1: only the adjacent strips' tails matter (see rf_xchg_put_part) */
int rf_plan_shard_resolve_lines(rf_plan* plan, const void* gathered_tails_dev, int nshards, int64_t nlines,
                                void* ext_all_dev, void* stream);
int rf_plan_stage2_ext(rf_plan* plan, const void* in_dev, void* out_dev, const void* ext_dev, void* stream);
int rf_plan_stage1(rf_plan* plan, const void* in_dev, void* out_dev, void* tails_dev, void* stream);
int rf_plan_stage2(rf_plan* plan, const void* in_dev, void* out_dev,
                   const void* gathered_tails_dev, int nshards, int shard_rank, void* stream);

/*
 * One filter over several GPUs of ONE process (what RecFilter::realize / profile use with RECFILTER_GPUS=n; no
 * reference equivalent: lib/recfilter.cpp:932-1016 drives one device).  The outermost dimension is cut into n equal
 * parts: independent parts when it carries no scans (stacks of images, audio channels -- nothing is exchanged),
 * strips otherwise (rf_plan_stage1, the order-r boundary tails pulled from the peers with cudaMemcpyPeerAsync over
 * NVLink, rf_plan_stage2).  Host buffers are dense, the whole array.
 */
typedef struct rf_mgpu rf_mgpu;
int  rf_mgpu_create(const rf_desc* desc, int ngpus, rf_mgpu** out);
void rf_mgpu_destroy(rf_mgpu* m);
int  rf_mgpu_ngpus(const rf_mgpu* m);
int  rf_mgpu_describe(const rf_mgpu* m, char* buf, size_t n);
int  rf_mgpu_execute_host(rf_mgpu* m, const void* in_host, void* out_host);           /* H2D + filter + D2H, synchronous */
int  rf_mgpu_profile(rf_mgpu* m, const void* in_host, int iters, float* ms_per_iter); /* device-resident, slowest GPU */
const char* rf_mgpu_last_error(void);

/*
 * Exchange windows for the strip tails (peer to peer over NVLink; no collective library, no host in the data
 * path).  Every rank owns a window of two generations of [nranks][bytes_per_rank] in its device memory.
 *   rf_xchg_put   copies this rank's tails into its slot of EVERY rank's window and then raises its arrival word
 *                 there (stream ordered, asynchronous);
 *   rf_xchg_wait  makes `stream` wait until all nranks slots of the current step have arrived and returns the
 *                 gathered [nranks][bytes_per_rank] array (what rf_plan_stage2 takes).
 * One process per GPU: pass the 64-byte rf_xchg_ipc_handle of every rank to rf_xchg_open_peer of every other
 * rank once (any host channel).  One process, several GPUs: rf_xchg_set_peer.  The wait gives up after a bounded
 * number of polls instead of hanging the device (rf_xchg_check reports it).
 */
typedef struct rf_xchg rf_xchg;
int    rf_xchg_create(size_t bytes_per_rank, int nranks, int rank, rf_xchg** out);   /* on the current device */
void   rf_xchg_destroy(rf_xchg* x);
size_t rf_xchg_handle_bytes(void);
int    rf_xchg_ipc_handle(rf_xchg* x, void* handle);
int    rf_xchg_open_peer(rf_xchg* x, int peer, const void* handle);
int    rf_xchg_set_peer(rf_xchg* x, int peer, rf_xchg* peer_window);
int    rf_xchg_put(rf_xchg* x, const void* src_dev, size_t bytes, void* stream);
int    rf_xchg_wait(rf_xchg* x, void* stream, void** gathered_dev);
/*
 * Neighbour exchange.  When rf_plan_shard_neighbors_suffice(plan) is 1 -- the filter forgets what entered a strip
 * before it leaves it, i.e. every entry of the whole-strip transition matrices is below 1e-30 -- the carries entering
 * strip s follow from the tails of strips s-1 and s+1 (and its own); the tails of farther strips may be left zero.
 * rf_xchg_put_part assembles a step from parts: bytes [offset, offset + bytes) of the tails go into this rank's slot of
 * the windows of the ranks in peer_mask (bit p = rank p, the own bit = the own window), the part with last != 0 raises
 * the arrival word in every window touched; rf_xchg_wait_from waits for the ranks in from_mask only.  Causal scans'
 * vectors go to rank + 1, anticausal ones to rank - 1 (layout of the tails: [vectors][lines], rf_plan_shard_vectors).
 */
int    rf_xchg_put_part(rf_xchg* x, unsigned peer_mask, const void* src_dev, size_t offset, size_t bytes, int last, void* stream);
int    rf_xchg_wait_from(rf_xchg* x, unsigned from_mask, void* stream, void** gathered_dev);
int    rf_xchg_check(rf_xchg* x);
const char* rf_xchg_last_error(void);

/*
 * Device-side stopwatch for a sequence of asynchronous calls on one stream (CUDA events):
 * what RecFilter::profile needs to time a chain of plans (lib/recfilter.cpp:998-1011, which uses
 * an unsynchronised wall clock).  rf_clock_end waits for the work, returns the elapsed
 * milliseconds and releases the clock.
 */
int rf_clock_begin(void* stream, void** clock);
int rf_clock_end(void* clock, void* stream, float* ms);

/* device memory helpers so that non-CUDA hosts (ctypes, cgo, JNI) can stage buffers */
int rf_malloc(void** dev_ptr, size_t bytes);
int rf_free(void* dev_ptr);
int rf_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
int rf_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
int rf_memset(void* dst_dev, int value, size_t bytes);
int rf_malloc_host(void** host_ptr, size_t bytes);   /* pinned */
int rf_free_host(void* host_ptr);
int rf_synchronize(void);

#ifdef __cplusplus
}
#endif
#endif /* RECFILTER_B200_H_ */
