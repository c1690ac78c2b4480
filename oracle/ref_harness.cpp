/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Thin extern "C" surface over the UNMODIFIED reference's CPU implementation of
 * the hot path (the "LLVM backend" ThreadState, plain C++ on the nanothread
 * pool: /root/reference/src/llvm_ts.cpp:265 block_reduce, :352
 * block_prefix_reduce, :479 reduce_dot, :706 compress, :785 block_mkperm).
 * Linked by oracle/Makefile against objects compiled from the reference sources
 * where they lie; the result lives in oracle/_ref/libref_cpu.so.
 *
 * jit_init(LLVM) cannot succeed in this image (no libLLVM with the LLVM-C API,
 * /root/reference/src/llvm_api.cpp:61-89), but none of the five primitives
 * touches LLVM, so -- as verified in SURVEY.md D8 -- the harness flips the
 * backend bit itself and then calls the public jit_* entry points unchanged.
 */
#include <drjit-core/jit.h>
#include <nanothread/nanothread.h>
#include "src/internal.h"
#include "src/llvm.h"
#include "src/util.h"
#include <cstring>
#include <stdexcept>

// Stand-in for the LZ4 dictionary that the reference's build embeds with
// cmake/bin2c.cmake (disk kernel cache only; never read on this path).
extern "C" {
__attribute__((visibility("default"))) extern const char kernels_dict[1];
const char kernels_dict[1] = { 0 };
}

static char ref_error[512];

#define REF_GUARD(stmt)                                                        \
    try { stmt; return 0; }                                                    \
    catch (const std::exception &e) {                                          \
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1); return 1;         \
    }

extern "C" {

__attribute__((visibility("default"))) const char *ref_last_error() { return ref_error; }

__attribute__((visibility("default"))) int ref_init(uint32_t threads) {
    try {
        jit_set_log_level_stderr(LogLevel::Error);
        jit_init(1u << (uint32_t) JitBackend::LLVM); // fails softly: no libLLVM
        {
            lock_guard guard(state.lock);
            state.backends |= 1u << (uint32_t) JitBackend::LLVM;
            if (jitc_llvm_vector_width == 0)
                jitc_llvm_vector_width = 8;
        }
        if (threads)
            jit_llvm_set_thread_count(threads);
        return 0;
    } catch (const std::exception &e) {
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1);
        return 1;
    }
}

__attribute__((visibility("default"))) uint32_t ref_pool_size() { return pool_size(nullptr); }

__attribute__((visibility("default"))) void ref_sync() { jit_sync_thread(); }

__attribute__((visibility("default"))) int ref_block_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                     const void *in, void *out) {
    REF_GUARD(jit_block_reduce(JitBackend::LLVM, (VarType) vt, (ReduceOp) op,
                               size, block_size, in, out);
              jit_sync_thread());
}

// positional contract of the reference: (.., SIZE, BLOCK_SIZE, ..), see
// /root/reference/src/api.cpp:1331-1337 and src/util.cpp:55-61
__attribute__((visibility("default"))) int ref_block_prefix_reduce(int vt, int op, uint32_t size,
                            uint32_t block_size, int exclusive, int reverse,
                            const void *in, void *out) {
    REF_GUARD(jit_block_prefix_reduce(JitBackend::LLVM, (VarType) vt,
                                      (ReduceOp) op, size, block_size,
                                      exclusive, reverse, in, out);
              jit_sync_thread());
}

__attribute__((visibility("default"))) int ref_reduce_dot(int vt, const void *a, const void *b, uint32_t size,
                   void *out) {
    REF_GUARD({
        lock_guard guard(state.lock);
        jitc_reduce_dot(JitBackend::LLVM, (VarType) vt, a, b, size, out);
    } jit_sync_thread());
}

__attribute__((visibility("default"))) int ref_compress(const uint8_t *in, uint32_t size, uint32_t *out,
                 uint32_t *count) {
    REF_GUARD(*count = jit_compress(JitBackend::LLVM, in, size, out));
}

__attribute__((visibility("default"))) int ref_block_mkperm(const uint32_t *values, uint32_t size,
                     uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                     uint32_t *offsets, uint32_t *unique) {
    REF_GUARD(*unique = jit_block_mkperm(JitBackend::LLVM, values, size,
                                         block_size, bucket_count, perm,
                                         offsets);
              jit_sync_thread());
}

__attribute__((visibility("default"))) uint64_t ref_reduce_identity(int vt, int op) {
    return jit_reduce_identity((VarType) vt, (ReduceOp) op);
}

// all()/any() of a bool array: /root/reference/src/util.cpp:172-211.
// NOTE: overwrites up to 3 bytes past 'size' (src/init.cpp:919-939).
__attribute__((visibility("default"))) int ref_all(uint8_t *values, uint32_t size, int *result) {
    REF_GUARD({
        lock_guard guard(state.lock);
        *result = jitc_all(JitBackend::LLVM, values, size) ? 1 : 0;
    });
}

__attribute__((visibility("default"))) int ref_any(uint8_t *values, uint32_t size, int *result) {
    REF_GUARD({
        lock_guard guard(state.lock);
        *result = jitc_any(JitBackend::LLVM, values, size) ? 1 : 0;
    });
}

} // extern "C"
