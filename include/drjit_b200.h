/*
 * include/drjit_b200.h -- C-ABI of the B200-native data-parallel primitive
 * library that stands in for drjit-core's CUDA primitive path.
 *
 * This is the "tier 1" boundary: plain C, plain pointers and sizes, no C++ or
 * torch types.  Every entry point names the reference interface it replaces
 * (paths relative to the drjit-core tree).  The jit.h-signature layer that
 * existing C++ callers link against sits on top of it in
 * include/drjit_b200_jit.h.
 *
 * Conventions
 *   - 'stream' is a cudaStream_t / CUstream passed as void*.  NULL selects the
 *     library's own per-device stream (the reference keeps exactly one stream
 *     per device, src/cuda_core.cpp:480), see b200_stream().
 *   - All primitives are asynchronous on 'stream' unless stated otherwise;
 *     temporaries come from the CUDA stream-ordered allocator.
 *   - 'vt' / 'op' / 'mode' use the reference enum values: VarType
 *     (include/drjit-core/jit.h:597-611), ReduceOp (:990-1014), ReduceMode
 *     (:1017-1066).
 *   - Return value 0 = success.  Non-zero: B200_ERR_* below; the message is
 *     available from b200_last_error() (thread-local).  Argument errors are the
 *     ones the reference raises as std::runtime_error.
 *   - There is no CPU fallback: without a CUDA device every call fails with
 *     B200_ERR_CUDA.
 */
#ifndef DRJIT_B200_H
#define DRJIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#  define B200_API __declspec(dllexport)
#else
#  define B200_API __attribute__((visibility("default")))
#endif

enum {
    B200_OK = 0,
    B200_ERR_INVALID = 1,     /* reference: jitc_raise (bad size / block size) */
    B200_ERR_UNSUPPORTED = 2, /* reference: "no existing kernel for type=.., op=.." */
    B200_ERR_CUDA = 3,        /* reference: cuda_check -> jitc_fail */
    B200_ERR_SYNC_FORBIDDEN = 4 /* reference: jitc_sync_thread under JitFlag::ForbidSynchronization,
                                   src/init.cpp:503-505 */
};

/* JitFlag bits honoured on this path (jit.h:1734-1742, same values) */
enum {
    B200_FLAG_KERNEL_HISTORY = 1 << 15,        /* record every primitive call (b200_kernel_history) */
    B200_FLAG_LAUNCH_BLOCKING = 1 << 16,       /* synchronise after every kernel launch */
    B200_FLAG_FORBID_SYNCHRONIZATION = 1 << 17 /* blocking entry points fail with B200_ERR_SYNC_FORBIDDEN */
};

/* KernelType values of the kernel history (jit.h:2597-2634) */
enum {
    B200_KERNEL_BLOCK_REDUCE = 1, B200_KERNEL_BLOCK_PREFIX_REDUCE = 2, B200_KERNEL_DOT = 3,
    B200_KERNEL_COMPRESS = 5, B200_KERNEL_MKPERM = 6, B200_KERNEL_MEMCPY = 7,
    B200_KERNEL_MEMSET = 8,
    B200_KERNEL_SCATTER = 64, /* stand-alone scatter kernels: fused into JIT kernels in the reference */
    B200_KERNEL_GATHER = 65   /* stand-alone packet gather */
};

/* VarType values used on this path (jit.h:597-611) */
enum {
    B200_VT_BOOL = 1, B200_VT_INT8 = 3, B200_VT_UINT8 = 4, B200_VT_INT16 = 5,
    B200_VT_UINT16 = 6, B200_VT_INT32 = 7, B200_VT_UINT32 = 8, B200_VT_INT64 = 9,
    B200_VT_UINT64 = 10, B200_VT_POINTER = 11, B200_VT_FLOAT16 = 13,
    B200_VT_FLOAT32 = 14, B200_VT_FLOAT64 = 15
};

/* ReduceOp (jit.h:990-1014) */
enum {
    B200_OP_IDENTITY = 0, B200_OP_ADD = 1, B200_OP_MUL = 2, B200_OP_MIN = 3,
    B200_OP_MAX = 4, B200_OP_AND = 5, B200_OP_OR = 6
};

/* ReduceMode (jit.h:1017-1066); Expand / Permute are compiler-level strategies
 * with no kernel of their own and are rejected here. */
enum {
    B200_MODE_AUTO = 0, B200_MODE_DIRECT = 1, B200_MODE_LOCAL = 2,
    B200_MODE_NO_CONFLICTS = 3
};

/* ---------------------------------------------------------------- runtime */

/* Replaces jit_init(1 << CUDA) / jitc_cuda_init (src/cuda_core.cpp:266-539) for
 * this path: enumerates devices, creates one stream per device, loads the
 * sm_100a kernels linked into this library (instead of decompressing and
 * JIT-compiling resources/kernels_75.lz4).  Idempotent. */
B200_API int b200_init(void);
B200_API int b200_shutdown(void);
B200_API const char *b200_last_error(void);
B200_API int b200_device_count(void);            /* jit_cuda_device_count, jit.h */
B200_API int b200_set_device(int device);        /* jit_cuda_set_device, src/init.cpp:468-496 */
B200_API int b200_device(void);
B200_API void *b200_stream(void);                /* jit_cuda_stream, jit.h:192 */
B200_API int b200_sync(void *stream);            /* jit_sync_thread, src/init.cpp:499+ */
B200_API int b200_sync_stream(void *other);      /* jit_cuda_sync_stream, jit.h:243-255: 'other' waits for
                                                    the work enqueued on the library stream so far */
B200_API int b200_sm_count(void);

/* jit_set_flags / jit_flags / jit_set_flag (jit.h:1815-1824) for the JitFlag bits
 * above; other bits are stored and reported back but have no effect here. */
B200_API void b200_set_flags(uint32_t flags);
B200_API uint32_t b200_flags(void);
B200_API void b200_set_flag(uint32_t flag, int enable);

/* jit_kernel_history / jit_kernel_history_clear (jit.h:2636-2742): with
 * B200_FLAG_KERNEL_HISTORY set every primitive call is bracketed by two events on
 * its stream (CUDAThreadState::submit does that per launch, src/cuda_ts.cpp:23-46).
 * Waits for the recorded calls, writes up to 'capacity' records, clears the
 * history and returns the number of records there were. */
typedef struct B200KernelRecord {
    int type;                 /* B200_KERNEL_* (KernelType) */
    uint64_t size;            /* KernelHistoryEntry::size */
    float execution_time_ms;  /* KernelHistoryEntry::execution_time */
} B200KernelRecord;
B200_API int b200_kernel_history(B200KernelRecord *out, int capacity);
B200_API void b200_kernel_history_clear(void);

/* jit_malloc / jit_free (src/malloc.cpp:102-306): kind 0 = device, 1 = pinned
 * host.  Sizes are rounded like the reference (>= 64 B, next power of two,
 * src/malloc.cpp:113-124) so code that relied on that padding keeps working. */
B200_API void *b200_malloc(size_t size, int kind);
/* Released in the order of the stream (and on the device) the block was allocated
 * on.  A block that was last used on ANOTHER stream is released with
 * b200_free_on(that stream, ptr). */
B200_API int b200_free(void *ptr);
B200_API int b200_free_on(void *stream, void *ptr);
/* jit_malloc_migrate (jit.h:495-518) between kind 0 (device) and 1 (pinned host);
 * asynchronous on the library stream. */
B200_API void *b200_malloc_migrate(void *ptr, int kind, int move);
B200_API int b200_memcpy(void *dst, const void *src, size_t size);                       /* jit_memcpy, jit.h:2202 */
B200_API int b200_memcpy_async(void *stream, void *dst, const void *src, size_t size);   /* jit_memcpy_async, jit.h:2205 */

/* jit_memset_async (jit.h:2199; CUDAThreadState::memset_async,
 * src/cuda_ts.cpp:129-183 incl. the fill_64 kernel of resources/misc.cuh:28-32):
 * writes 'size' elements of 'isize' in {1,2,4,8} bytes taken from host pointer
 * 'src'. */
B200_API int b200_memset_async(void *stream, void *ptr, uint64_t size, uint32_t isize,
                               const void *src);

/* jit_reduce_identity (jit.h:2840; src/var.cpp:2642-2652) */
B200_API uint64_t b200_reduce_identity(int vt, int op);

/* ------------------------------------------------------------- reductions */

/* jit_block_reduce (jit.h:2239-2245) = CUDAThreadState::block_reduce
 * (src/cuda_ts.cpp:195-352) + kernels resources/block_reduce.cuh.
 * out[b] = op over in[b*block_size .. min((b+1)*block_size, size)).
 * Errors: block_size == 0 or > size -> B200_ERR_INVALID; unsupported
 * (type, op) -> B200_ERR_UNSUPPORTED; size == 0 is a no-op. */
B200_API int b200_block_reduce(void *stream, int vt, int op, uint64_t size,
                               uint64_t block_size, const void *in, void *out);

/* jit_reduce (jit.h:2219-2221) == block_reduce with block_size == size
 * (src/util.cpp:42-45) */
B200_API int b200_reduce(void *stream, int vt, int op, const void *in, uint64_t size,
                         void *out);

/* jitc_reduce_dot (src/util.cpp:63-67) = CUDAThreadState::reduce_dot
 * (src/cuda_ts.cpp:354-398) + resources/reduce_2.cuh.  f16 / f32 / f64. */
B200_API int b200_reduce_dot(void *stream, int vt, const void *a, const void *b,
                             uint64_t size, void *out);

/* jitc_all / jitc_any (+ _async) (src/util.cpp:153-211,
 * ThreadState::block_reduce_bool src/init.cpp:919-939).  'out' is a device or
 * pinned byte.  Unlike the reference nothing is written past values[size). */
B200_API int b200_all_async(void *stream, const uint8_t *values, uint64_t size, uint8_t *out);
B200_API int b200_any_async(void *stream, const uint8_t *values, uint64_t size, uint8_t *out);
B200_API int b200_all(void *stream, const uint8_t *values, uint64_t size, int *result);  /* synchronises */
B200_API int b200_any(void *stream, const uint8_t *values, uint64_t size, int *result);  /* synchronises */

/* ------------------------------------------------------------ prefix scans */

/* jit_block_prefix_reduce (jit.h:2365-2373, positional order (size,
 * block_size), src/util.cpp:55-61) = CUDAThreadState::block_prefix_reduce
 * (src/cuda_ts.cpp:530-681) + resources/block_prefix_reduce.cuh.
 * in == out is allowed. */
B200_API int b200_block_prefix_reduce(void *stream, int vt, int op, uint64_t size,
                                      uint64_t block_size, int exclusive, int reverse,
                                      const void *in, void *out);

/* Extension used by the multi-GPU front end and by > 2^32-element arrays:
 * whole-array scan (one block) seeded with *carry_in (device scalar of the
 * value type, may be NULL) and reporting the inclusive total of
 * carry_in (+) array in *carry_out (device scalar, may be NULL). */
B200_API int b200_prefix_reduce_carry(void *stream, int vt, int op, uint64_t size,
                                      int exclusive, int reverse, const void *in,
                                      void *out, const void *carry_in, void *carry_out);

/* Whole-array prefix reduction whose TILE prefixes are supplied by the caller:
 * tile_seeds[t] (device, value type) = reduction of everything in front of tile t
 * (behind it when reverse), tiles of b200_scan_tile_elems(vt) elements counted from
 * the start of the array.  No look-back chain: every tile is independent, which
 * is what the sharded front end wants -- it reduces its shard once anyway (for the
 * exchange of the rank totals) and gets the tile sums from that pass.  New entry
 * point (the reference has no counterpart).  Requires 16-byte aligned arrays. */
B200_API uint32_t b200_scan_tile_elems(int vt);
B200_API int b200_prefix_reduce_seeded(void *stream, int vt, int op, uint64_t size, int exclusive,
                                       int reverse, const void *in, void *out,
                                       const void *tile_seeds);

/* ---------------------------------------------------------------- compress */

/* jit_compress (jit.h:2387) = CUDAThreadState::compress
 * (src/cuda_ts.cpp:683-763) + resources/compress.cuh.  Synchronises the stream
 * and returns the count through *count (host memory).  'in' is never written
 * (the reference zero-fills its trailer). */
B200_API int b200_compress(void *stream, const uint8_t *in, uint64_t size, uint32_t *out,
                           uint32_t *count);
/* Asynchronous form: *count_dev is device or pinned memory. */
B200_API int b200_compress_async(void *stream, const uint8_t *in, uint64_t size,
                                 uint32_t *out, uint32_t *count_dev);

/* ------------------------------------------------------------------ mkperm */

/* jit_block_mkperm (jit.h:2426-2432) = CUDAThreadState::block_mkperm
 * (src/cuda_ts.cpp:788-975) + resources/mkperm.cuh.  'offsets' (may be NULL) is
 * host-accessible memory of (4 * bucket_count + 1) u32 receiving
 * (id, start, size, 0) records of the non-empty buckets -- in ascending id
 * order, like the reference's CPU path (src/llvm_ts.cpp:871-887) -- and the
 * unique count at [4 * bucket_count]; only when block_size == size.  The
 * permutation is stable for every bucket count.  *unique receives the return
 * value of the reference call.  Synchronises when offsets != NULL. */
B200_API int b200_block_mkperm(void *stream, const uint32_t *values, uint32_t size,
                               uint32_t block_size, uint32_t bucket_count,
                               uint32_t *perm, uint32_t *offsets, uint32_t *unique);

/* Asynchronous form: everything is enqueued (including the copy of the records
 * into 'offsets'), nothing is waited for; after the stream has been synchronised
 * offsets[4 * bucket_count] holds the return value of the reference call.  This is
 * what a host that holds a lock while enqueueing and drops it while waiting
 * (state.lock / unlock_guard, src/cuda_ts.cpp:964-967) calls. */
B200_API int b200_block_mkperm_async(void *stream, const uint32_t *values, uint32_t size,
                                     uint32_t block_size, uint32_t bucket_count,
                                     uint32_t *perm, uint32_t *offsets);

/* The consumer of jit_block_mkperm in vectorised method dispatch, jitc_var_call_reduce
 * (jit.h:2510, src/call.cpp:1268-1389), up to the point where it creates variables:
 * callable ids in [0, id_bound] (0 = the null callable; the reference adds that bucket
 * itself, call.cpp:1292) are grouped by jit_block_mkperm(size, size, id_bound + 1) and the
 * (id, start, size, 0) records of the non-empty buckets arrive in `offsets` (host
 * accessible, 4 * (id_bound + 1) + 1 words) ALREADY ORDERED BY SIZE, largest first
 * (ties: ascending id) -- the order the reference establishes with a host-side std::sort
 * after its event wait (call.cpp:1346-1352; std::sort leaves the order of ties open).
 * offsets[4 * (id_bound + 1)] and *unique = number of records.  perm[start .. start + size)
 * of a record are the (stable) element indices of that callable. */
B200_API int b200_call_reduce(void *stream, const uint32_t *ids, uint32_t size, uint32_t id_bound,
                              uint32_t *perm, uint32_t *offsets, uint32_t *unique);
B200_API int b200_call_reduce_async(void *stream, const uint32_t *ids, uint32_t size,
                                    uint32_t id_bound, uint32_t *perm, uint32_t *offsets);

/* Phase 1 of the above on its own (per-bucket counts of the whole array into
 * device memory hist[bucket_count]); the multi-GPU front end all-reduces it. */
B200_API int b200_mkperm_histogram(void *stream, const uint32_t *values, uint64_t size,
                                   uint32_t bucket_count, uint32_t *hist);

/* ----------------------------------------------------------------- scatter */

/* Stand-alone form of the scatter-reduce that the reference splices into fused
 * JIT kernels (jitc_cuda_render_scatter_reduce, src/cuda_scatter.cpp:246-354;
 * warp pre-reduction :125-244; op built by jitc_var_scatter src/op.cpp:2899-3086):
 *   if (mask == NULL || mask[i]) target[index[i]] op= value[i],  i < n.
 * op in {Add, Min, Max, And, Or}; legal (type, op) pairs as
 * jitc_can_scatter_reduce (src/op.cpp:2735-2820).  mode: Auto -> Local (the
 * reference's default flag set, jit.h:1766-1774). */
B200_API int b200_scatter_reduce(void *stream, int vt, int op, int mode, void *target,
                                 const void *value, const uint32_t *index,
                                 const uint8_t *mask, uint64_t n);

/* jit_can_scatter_reduce (jit.h:1123) for the CUDA backend on sm_100 */
B200_API int b200_can_scatter_reduce(int vt, int op);

/* Stand-alone form of jit_var_scatter_inc (jit.h:1125-1143; emitter
 * jitc_cuda_render_scatter_inc, src/cuda_scatter.cpp:356-393):
 *   if (mask == NULL || mask[i]) { out[i] = target[index[i]]; target[index[i]] += 1; }
 * atomically and warp aggregated; masked entries receive 0.  Which of several
 * entries with the same index gets which old value is unspecified (as in the
 * reference); the values handed out for a counter are consecutive. */
B200_API int b200_scatter_inc(void *stream, uint32_t *target, const uint32_t *index,
                              const uint8_t *mask, uint32_t *out, uint64_t n);

/* Stand-alone form of the reducing packet scatter (jit_var_scatter_packet,
 * jit.h:1117; emitter jitc_cuda_render_scatter_reduce_packet,
 * src/cuda_packet.cpp:169-327):
 *   if (mask == NULL || mask[i]) target[index[i] * width + k] op= values[k][i], k < width
 * width in {1, 2, 4, 8}; values: HOST array of `width` device pointers (the
 * reference takes `width` separate variables).  f16 / f32 / f64 Add, f32 Min / Max,
 * u32 / i32 integer operations.  f32 Add issues red.global.add.v2/.v4.f32, f16 Add
 * red.global.v2/.v4/.v8.f16.add.noftz (src/cuda_packet.cpp:229-266); those need the
 * target aligned to min(width * sizeof(T), 16) bytes. */
B200_API int b200_scatter_reduce_packet(void *stream, int vt, int op, int mode, void *target,
                                        const void *const *values, uint32_t width,
                                        const uint32_t *index, const uint8_t *mask, uint64_t n);

/* b200_scatter_reduce with the other index types the reference accepts (int32, uint64,
 * int64; jitc_var_scatter, src/op.cpp:2899-3086) and with op == Identity: the plain
 * scatter target[index[i]] = value[i] of any 1 / 2 / 4 / 8 byte type (duplicate
 * indices: one of the values wins).  index_vt: VarType of the index array. */
B200_API int b200_scatter_reduce_idx(void *stream, int vt, int op, int mode, void *target,
                                     const void *value, const void *index, int index_vt,
                                     const uint8_t *mask, uint64_t n);

/* jit_var_scatter_packet WITHOUT reduction (jit.h:1117 with ReduceOp::Identity; emitter
 * jitc_cuda_render_scatter_packet, src/cuda_packet.cpp:329-443):
 *   if (mask == NULL || mask[i]) target[index[i] * width + k] = values[k][i], k < width
 * for any 1 / 2 / 4 / 8 byte type, width in {1, 2, 4, 8}; a packet leaves in vector
 * stores of up to 128 bits. */
B200_API int b200_scatter_packet(void *stream, int vt, void *target, const void *const *values,
                                 uint32_t width, const uint32_t *index, const uint8_t *mask,
                                 uint64_t n);

/* Stand-alone form of jit_var_gather_packet (jit.h:981; emitter
 * jitc_cuda_render_gather_packet, src/cuda_packet.cpp:18-166): array-of-structures ->
 * structure-of-arrays,
 *   out[k][i] = (mask == NULL || mask[i]) ? source[index[i] * width + k] : 0,  k < width
 * with vector loads of up to 128 bits.  out: HOST array of `width` device pointers. */
B200_API int b200_gather_packet(void *stream, int vt, const void *source, void *const *out,
                                uint32_t width, const uint32_t *index, const uint8_t *mask,
                                uint64_t n);

/* --------------------------------------------------------------- multi-GPU */

/* Large reductions, whole-array prefix reductions and the mkperm histogram sharded
 * over the GPUs of one box (BASELINE.json north_star; the reference has no
 * multi-GPU code on this path).  One process per GPU; rank r owns a contiguous
 * shard.  The W ranks exchange their totals through PEER-MAPPED MAILBOXES over
 * NVLink -- no communication-library call on the data path:
 *   b200_sharded_create    allocates this rank's mailbox (on the current device);
 *   b200_sharded_export    writes its CUDA IPC handle (b200_sharded_handle_bytes());
 *   b200_sharded_connect   maps the mailboxes of all ranks: `handles` holds the W
 *                          exported handles in rank order (gathered by the host
 *                          program over any channel, e.g. torch.distributed).
 * Collectives (every rank calls them in the same order, each rank on ONE stream):
 * all are asynchronous; a rank that waits more than 20 s for a peer gives up and
 * the next call on the context fails. */
typedef struct B200Sharded B200Sharded;
B200_API int b200_sharded_create(int rank, int world, B200Sharded **ctx);
B200_API int b200_sharded_handle_bytes(void);
B200_API int b200_sharded_export(B200Sharded *ctx, void *handle);
B200_API int b200_sharded_connect(B200Sharded *ctx, const void *handles);
B200_API int b200_sharded_destroy(B200Sharded *ctx);

/* Reduction of the global array; every rank receives the result in *out (device).
 * The W partials are combined in rank order on every rank: bit-exact for integers,
 * one fixed order for floating point. */
B200_API int b200_sharded_reduce(B200Sharded *ctx, void *stream, int vt, int op, const void *in,
                                 uint64_t local_size, void *out);

/* Dot product of two equally sharded arrays ("reduce / block_reduce / dot"):
 * jitc_reduce_dot (src/util.cpp:63-67) on the shard, then the exchange of
 * b200_sharded_reduce with ReduceOp::Add; every rank receives the result.  f16 / f32 / f64. */
B200_API int b200_sharded_reduce_dot(B200Sharded *ctx, void *stream, int vt, const void *a, const void *b,
                                     uint64_t local_size, void *out);

/* Prefix reduction of the global array (block_size == global size); rank r's shard
 * of the result goes to `out`.  16-byte aligned shards of fewer than 2^32 - 2^14
 * elements; no float16.  Three launches per rank: tile sums of the shard, one CTA
 * that exchanges the shard totals and turns the tile sums into tile prefixes, the
 * seeded streaming scan (b200_prefix_reduce_seeded). */
B200_API int b200_sharded_prefix_reduce(B200Sharded *ctx, void *stream, int vt, int op,
                                        uint64_t local_size, int exclusive, int reverse,
                                        const void *in, void *out);

/* The same prefix reduction over a BLOCK-CYCLIC layout (opt-in): the global array is cut
 * into blocks of `block_size` elements (a power of two, at least one 32 KiB scan tile),
 * global block b lives on rank b % world as local block b / world; local_size: elements
 * of this rank (whole blocks, the same on every rank).  ONE pass over the data (8 B /
 * element per GPU, against 12 for contiguous shards, where a rank cannot start before it
 * knows the totals of all ranks in front of it): every rank runs one chained streaming
 * scan, block totals travel as single 16-byte stores into peer-mapped tables.  Forward
 * only, no float16. */
B200_API int b200_sharded_prefix_reduce_cyclic(B200Sharded *ctx, void *stream, int vt, int op,
                                               uint64_t local_size, uint64_t block_size,
                                               int exclusive, const void *in, void *out);

/* Global bucket counts of mkperm keys (phase 1 of jit_block_mkperm over the sharded
 * array): hist[bucket_count] on every rank; `before` (may be NULL) receives the
 * counts of the ranks in front of this one, i.e. this rank's first output slot
 * inside every bucket.  bucket_count <= 65536. */
B200_API int b200_sharded_histogram(B200Sharded *ctx, void *stream, const uint32_t *values,
                                    uint64_t local_size, uint32_t bucket_count, uint32_t *hist,
                                    uint32_t *before);

/* --------------------------------------------------------------- telemetry */

/* Number of kernels this library has launched since load (all threads). */
B200_API uint64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
