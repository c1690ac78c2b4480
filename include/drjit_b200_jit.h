/*
 * include/drjit_b200_jit.h -- jit.h-signature layer ("tier 2") over the C-ABI
 * in drjit_b200.h.
 *
 * drjit-core's public API has C++ linkage (include/drjit-core/jit.h has no
 * extern "C"; with a C++ compiler its enums are `enum class`,
 * include/drjit-core/macros.h:37-53), so existing callers bind to MANGLED
 * names such as _Z16jit_block_reduce10JitBackend7VarType8ReduceOpjjPKvPv.
 * libdrjit_core_b200.so exports exactly those names for the entry points of
 * the data-parallel primitive path, with the reference's argument order, enum
 * values and error behaviour (invalid arguments throw std::runtime_error, as
 * jitc_raise does, src/log.cpp:165-169).  A translation unit compiled against
 * the reference's own jit.h therefore links against this library unchanged for
 * these symbols.
 *
 * Only JitBackend::CUDA is served.  Any other backend throws: this library has
 * no CPU fallback.
 *
 * The declarations below restate the ABI (names, parameter types, enumerator
 * values); each cites the jit.h line of the declaration it mirrors.
 */
#pragma once

#include <cstddef>
#include <cstdint>

#if !defined(DRJIT_B200_EXPORT)
#  define DRJIT_B200_EXPORT __attribute__((visibility("default")))
#endif

/* jit.h:47-61 */
enum class JitBackend : uint32_t { None = 0, CUDA = 1, LLVM = 2, Metal = 3 };

/* jit.h:597-611 */
enum class VarType : uint32_t {
    Void, Bool, BaseInt, Int8, UInt8, Int16, UInt16, Int32, UInt32, Int64, UInt64,
    Pointer, BaseFloat, Float16, Float32, Float64, Count
};

/* jit.h:990-1014 */
enum class ReduceOp : uint32_t { Identity, Add, Mul, Min, Max, And, Or, Count };

/* jit.h:1017-1066 */
enum class ReduceMode : uint32_t { Auto, Direct, Local, NoConflicts, Expand, Permute };

/* jit.h:1734-1742 (the JitFlag bits this path honours; the other bits are stored
 * and reported back by jit_flags() without effect) */
enum class JitFlag : uint32_t {
    KernelHistory = 1 << 15, LaunchBlocking = 1 << 16, ForbidSynchronization = 1 << 17
};

/* jit.h:2597-2634 */
enum KernelType : uint32_t {
    JIT, BlockReduce, BlockPrefixReduce, Dot, BatchedGemm, Compress, MkPerm, Memcpy, Memset,
    Poke, Aggregate, LLVMHostFunc
};

/* jit.h:2636-2650 */
enum KernelRecordingMode : uint32_t { Inactive, Recorded, Replayed };

/* jit.h:2652-2709 (same layout) */
struct KernelHistoryEntry {
    JitBackend backend;
    KernelType type;
    KernelRecordingMode recording_mode;
    uint64_t hash[2];
    char *ir;
    int uses_optix;
    int cache_hit;
    int cache_disk;
    uint32_t size;
    uint32_t input_count;
    uint32_t output_count;
    uint32_t operation_count;
    float codegen_time;
    float backend_time;
    float execution_time;
    void *event_start, *event_end;
    void *task;
};

/* ---- runtime ------------------------------------------------------------ */
extern DRJIT_B200_EXPORT void jit_cuda_sync_stream(uintptr_t stream);      /* jit.h:255 */
extern DRJIT_B200_EXPORT void jit_set_flags(uint32_t flags);               /* jit.h:1815 */
extern DRJIT_B200_EXPORT uint32_t jit_flags();                             /* jit.h:1818 */
extern DRJIT_B200_EXPORT void jit_set_flag(JitFlag flag, int enable);      /* jit.h:1821 */
extern DRJIT_B200_EXPORT int jit_flag(JitFlag flag);                       /* jit.h:1824 */
extern DRJIT_B200_EXPORT void jit_kernel_history_clear();                  /* jit.h:2712 */
extern DRJIT_B200_EXPORT KernelHistoryEntry *jit_kernel_history();         /* jit.h:2737 */
extern DRJIT_B200_EXPORT void *jit_malloc_migrate(void *ptr, JitBackend backend, int move); /* jit.h:516 */
extern DRJIT_B200_EXPORT void jit_init(uint32_t backends);                 /* jit.h:96 */
extern DRJIT_B200_EXPORT int jit_has_backend(JitBackend backend);          /* jit.h:119 */
extern DRJIT_B200_EXPORT void jit_shutdown(int light);                     /* jit.h:132 */
extern DRJIT_B200_EXPORT void jit_sync_thread();                           /* jit.h:141 */
extern DRJIT_B200_EXPORT int jit_cuda_device_count();                      /* jit.h:165 */
extern DRJIT_B200_EXPORT void jit_cuda_set_device(int device);             /* jit.h:175 */
extern DRJIT_B200_EXPORT int jit_cuda_device();                            /* jit.h:186 */
extern DRJIT_B200_EXPORT void *jit_cuda_stream();                          /* jit.h:192 */
extern DRJIT_B200_EXPORT void *jit_malloc(JitBackend backend, size_t size, int shared); /* jit.h:455 */
extern DRJIT_B200_EXPORT void jit_free(void *ptr);                         /* jit.h:472 */
extern DRJIT_B200_EXPORT void jit_memcpy(JitBackend backend, void *dst, const void *src, size_t size);       /* jit.h:2202 */
extern DRJIT_B200_EXPORT void jit_memcpy_async(JitBackend backend, void *dst, const void *src, size_t size); /* jit.h:2205 */
extern DRJIT_B200_EXPORT void jit_memset_async(JitBackend backend, void *ptr, uint32_t size,
                                               uint32_t isize, const void *src);                              /* jit.h:2199 */

/* ---- the primitives ------------------------------------------------------ */
extern DRJIT_B200_EXPORT uint64_t jit_reduce_identity(VarType vt, ReduceOp op);          /* jit.h:2840 */
extern DRJIT_B200_EXPORT int jit_can_scatter_reduce(JitBackend backend, VarType vt, ReduceOp op); /* jit.h:1123 */

/* jit.h:2219-2221 -- declared by the reference but never defined there with
 * this parameter order (src/api.cpp:1303 defines (.., size, in, out) as a local
 * symbol).  Both orders are provided. */
extern DRJIT_B200_EXPORT void jit_reduce(JitBackend backend, VarType type, ReduceOp op,
                                         const void *in, uint32_t size, void *out);
extern DRJIT_B200_EXPORT void jit_reduce(JitBackend backend, VarType type, ReduceOp op,
                                         uint32_t size, const void *in, void *out);

/* jit.h:2239-2245 */
extern DRJIT_B200_EXPORT void jit_block_reduce(JitBackend backend, VarType type, ReduceOp op,
                                               uint32_t size, uint32_t block_size,
                                               const void *in, void *out);

/* jit.h:2365-2373.  NOTE the positional contract is (size, block_size): the
 * header of the reference names the 4th/5th parameters the other way round but
 * src/api.cpp:1331-1337 -> src/util.cpp:55-61 bind them as below. */
extern DRJIT_B200_EXPORT void jit_block_prefix_reduce(JitBackend backend, VarType type, ReduceOp op,
                                                      uint32_t size, uint32_t block_size,
                                                      int exclusive, int reverse,
                                                      const void *in, void *out);

/* jit.h:2387 */
extern DRJIT_B200_EXPORT uint32_t jit_compress(JitBackend backend, const uint8_t *in, uint32_t size,
                                               uint32_t *out);

/* jit.h:2426-2432 */
extern DRJIT_B200_EXPORT uint32_t jit_block_mkperm(JitBackend backend, const uint32_t *values,
                                                   uint32_t size, uint32_t block_size,
                                                   uint32_t bucket_count, uint32_t *perm,
                                                   uint32_t *offsets);
