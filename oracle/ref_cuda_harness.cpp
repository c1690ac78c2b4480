/*
 * oracle/ref_cuda_harness.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Thin extern "C" surface over the UNMODIFIED reference built WITH its CUDA
 * backend (oracle/Makefile, target _ref/libref_cuda.so).  It lets the GPU
 * parity tests and bench.py run the reference's own CUDA kernels
 * (resources/kernels_75.lz4 -- compute_75 PTX that the driver JIT-compiles for
 * the B200) on the same device buffers as the sm_100a kernels of this
 * repository: CUDAThreadState::block_reduce / block_prefix_reduce / reduce_dot /
 * compress / block_mkperm (/root/reference/src/cuda_ts.cpp:195-975) through the
 * public jit_* entry points, and the scatter-reduce code generator
 * (/root/reference/src/cuda_scatter.cpp:246-354) through jit_var_mem_map +
 * jit_var_scatter + jit_eval.
 *
 * Both libraries use the device's primary context, so raw device pointers can
 * be shared.  The reference pads / overwrites memory past the end of compress
 * and all/any inputs (jit.h:2382-2383): callers pass buffers with slack.
 */
#include <drjit-core/jit.h>
#include "src/internal.h"
#include "src/util.h"
#include <cstring>
#include <stdexcept>

static char ref_error[1024];

#define REF_GUARD(stmt)                                                        \
    try { stmt; return 0; }                                                    \
    catch (const std::exception &e) {                                          \
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1); return 1;         \
    }

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char *refcuda_last_error() { return ref_error; }

/// Returns 0 when the CUDA backend came up
REF_API int refcuda_init() {
    try {
        jit_set_log_level_stderr(LogLevel::Error);
        jit_init(1u << (uint32_t) JitBackend::CUDA);
        if (!jit_has_backend(JitBackend::CUDA)) {
            strncpy(ref_error, "jit_init(): the CUDA backend is unavailable", sizeof(ref_error) - 1);
            return 1;
        }
        return 0;
    } catch (const std::exception &e) {
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1);
        return 1;
    }
}

REF_API void refcuda_shutdown() { jit_shutdown(1); }
REF_API void refcuda_sync() { jit_sync_thread(); }
REF_API void *refcuda_stream() { return jit_cuda_stream(); }

REF_API void *refcuda_malloc(size_t bytes) {
    try { return jit_malloc(JitBackend::CUDA, bytes, 0); } catch (...) { return nullptr; }
}
REF_API void refcuda_free(void *ptr) { jit_free(ptr); }
REF_API void *refcuda_malloc_pinned(size_t bytes) {
    try { return jit_malloc(JitBackend::None, bytes, 1); } catch (...) { return nullptr; }
}
REF_API int refcuda_memcpy(void *dst, const void *src, size_t bytes) {
    REF_GUARD(jit_memcpy(JitBackend::CUDA, dst, src, bytes));
}

REF_API int refcuda_block_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                                 const void *in, void *out) {
    REF_GUARD(jit_block_reduce(JitBackend::CUDA, (VarType) vt, (ReduceOp) op, size,
                               block_size, in, out));
}

// positional contract: (.., SIZE, BLOCK_SIZE, ..) -- src/api.cpp:1331-1337
REF_API int refcuda_block_prefix_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                                        int exclusive, int reverse, const void *in, void *out) {
    REF_GUARD(jit_block_prefix_reduce(JitBackend::CUDA, (VarType) vt, (ReduceOp) op, size,
                                      block_size, exclusive, reverse, in, out));
}

REF_API int refcuda_reduce_dot(int vt, const void *a, const void *b, uint32_t size, void *out) {
    // not exported at pointer level: same internal entry the variable API uses
    REF_GUARD({
        lock_guard guard(state.lock);
        jitc_reduce_dot(JitBackend::CUDA, (VarType) vt, a, b, size, out);
    });
}

REF_API int refcuda_compress(const uint8_t *in, uint32_t size, uint32_t *out, uint32_t *count) {
    REF_GUARD(*count = jit_compress(JitBackend::CUDA, in, size, out));
}

REF_API int refcuda_block_mkperm(const uint32_t *values, uint32_t size, uint32_t block_size,
                                 uint32_t bucket_count, uint32_t *perm, uint32_t *offsets,
                                 uint32_t *unique) {
    REF_GUARD(*unique = jit_block_mkperm(JitBackend::CUDA, values, size, block_size,
                                         bucket_count, perm, offsets));
}

REF_API int refcuda_memset_async(void *ptr, uint32_t size, uint32_t isize, const void *src) {
    REF_GUARD(jit_memset_async(JitBackend::CUDA, ptr, size, isize, src));
}

/// target[index[i]] op= value[i] (mask optional) through the reference's JIT:
/// the three arrays are mapped as evaluated variables, jit_var_scatter() records
/// the side effect and jit_eval() compiles + launches the fused kernel.  When
/// 'repeat' > 1 the scatter is recorded and evaluated that many times (timing).
REF_API int refcuda_scatter_reduce(int vt, int op, int mode, void *target, size_t target_size,
                                   const void *value, const uint32_t *index, const uint8_t *mask,
                                   size_t n, int repeat) {
    try {
        uint32_t t = jit_var_mem_map(JitBackend::CUDA, (VarType) vt, target, target_size, 0);
        uint32_t v = jit_var_mem_map(JitBackend::CUDA, (VarType) vt, (void *) value, n, 0);
        uint32_t i = jit_var_mem_map(JitBackend::CUDA, VarType::UInt32, (void *) index, n, 0);
        uint32_t m = mask ? jit_var_mem_map(JitBackend::CUDA, VarType::Bool, (void *) mask, n, 0)
                          : jit_var_bool(JitBackend::CUDA, true);
        for (int r = 0; r < (repeat < 1 ? 1 : repeat); ++r) {
            uint32_t t2 = jit_var_scatter(t, v, i, m, (ReduceOp) op, (ReduceMode) mode);
            jit_var_dec_ref(t);
            t = t2;
            jit_eval();
        }
        void *ptr = nullptr;
        jit_var_data(t, &ptr);
        if (ptr != target) // the JIT decided to work on a copy
            jit_memcpy(JitBackend::CUDA, target, ptr, target_size * jit_type_size((VarType) vt));
        jit_var_dec_ref(t);
        jit_var_dec_ref(v);
        jit_var_dec_ref(i);
        jit_var_dec_ref(m);
        return 0;
    } catch (const std::exception &e) {
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1);
        return 1;
    }
}

/// out[i] = target[index[i]]++ through the reference's JIT (jit_var_scatter_inc,
/// jit.h:1125-1143; PTX from src/cuda_scatter.cpp:356-393)
REF_API int refcuda_scatter_inc(uint32_t *target, size_t target_size, const uint32_t *index,
                                const uint8_t *mask, size_t n, uint32_t *out) {
    try {
        uint32_t t = jit_var_mem_map(JitBackend::CUDA, VarType::UInt32, target, target_size, 0);
        uint32_t i = jit_var_mem_map(JitBackend::CUDA, VarType::UInt32, (void *) index, n, 0);
        uint32_t m = mask ? jit_var_mem_map(JitBackend::CUDA, VarType::Bool, (void *) mask, n, 0)
                          : jit_var_bool(JitBackend::CUDA, true);
        uint32_t r = jit_var_scatter_inc(&t, i, m);
        jit_var_eval(r);
        jit_eval();
        void *ptr = nullptr;
        uint32_t r2 = jit_var_data(r, &ptr);
        jit_memcpy(JitBackend::CUDA, out, ptr, n * sizeof(uint32_t));
        jit_var_dec_ref(r2);
        uint32_t t2 = jit_var_data(t, &ptr);
        if (ptr != target) // the JIT decided to work on a copy
            jit_memcpy(JitBackend::CUDA, target, ptr, target_size * sizeof(uint32_t));
        jit_var_dec_ref(t2);
        jit_var_dec_ref(r);
        jit_var_dec_ref(t);
        jit_var_dec_ref(i);
        jit_var_dec_ref(m);
        return 0;
    } catch (const std::exception &e) {
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1);
        return 1;
    }
}

/// target[index[i] * width + k] op= values[k][i] through the reference's JIT
/// (jit_var_scatter_packet, jit.h:1117; PTX from src/cuda_packet.cpp:169-327)
REF_API int refcuda_scatter_packet(int vt, int op, int mode, void *target, size_t target_size,
                                   const void *const *values, size_t width, const uint32_t *index,
                                   const uint8_t *mask, size_t n) {
    try {
        uint32_t t = jit_var_mem_map(JitBackend::CUDA, (VarType) vt, target, target_size, 0);
        uint32_t vals[16];
        for (size_t k = 0; k < width; ++k)
            vals[k] = jit_var_mem_map(JitBackend::CUDA, (VarType) vt, (void *) values[k], n, 0);
        uint32_t i = jit_var_mem_map(JitBackend::CUDA, VarType::UInt32, (void *) index, n, 0);
        uint32_t m = mask ? jit_var_mem_map(JitBackend::CUDA, VarType::Bool, (void *) mask, n, 0)
                          : jit_var_bool(JitBackend::CUDA, true);
        uint32_t t2 = jit_var_scatter_packet(width, t, vals, i, m, (ReduceOp) op, (ReduceMode) mode);
        jit_var_dec_ref(t);
        t = t2;
        jit_eval();
        void *ptr = nullptr;
        uint32_t t3 = jit_var_data(t, &ptr);
        if (ptr != target)
            jit_memcpy(JitBackend::CUDA, target, ptr, target_size * jit_type_size((VarType) vt));
        jit_var_dec_ref(t3);
        jit_var_dec_ref(t);
        for (size_t k = 0; k < width; ++k)
            jit_var_dec_ref(vals[k]);
        jit_var_dec_ref(i);
        jit_var_dec_ref(m);
        return 0;
    } catch (const std::exception &e) {
        strncpy(ref_error, e.what(), sizeof(ref_error) - 1);
        return 1;
    }
}

REF_API int refcuda_can_scatter_reduce(int vt, int op) {
    return jit_can_scatter_reduce(JitBackend::CUDA, (VarType) vt, (ReduceOp) op);
}

// all()/any() of a bool array: src/util.cpp:153-211.  Overwrites up to 3 bytes
// past 'size' (src/init.cpp:919-939).
REF_API int refcuda_all(uint8_t *values, uint32_t size, int *result) {
    REF_GUARD({
        lock_guard guard(state.lock);
        *result = jitc_all(JitBackend::CUDA, values, size) ? 1 : 0;
    });
}

REF_API int refcuda_any(uint8_t *values, uint32_t size, int *result) {
    REF_GUARD({
        lock_guard guard(state.lock);
        *result = jitc_any(JitBackend::CUDA, values, size) ? 1 : 0;
    });
}
