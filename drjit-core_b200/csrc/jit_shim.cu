// jit_shim.cu -- the jit.h-signature layer declared in include/drjit_b200_jit.h.
//
// Mirrors src/api.cpp:1287-1351 (lock, forward to the backend's ThreadState)
// for the CUDA backend only: a global recursive lock around every call (the
// reference's state.lock), released while a call blocks on the stream
// (unlock_guard, src/cuda_ts.cpp:759, :964-967), errors re-raised as
// std::runtime_error like jitc_raise (src/log.cpp:165-169).
#include "../../include/drjit_b200_jit.h"
#include "../../include/drjit_b200.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

std::recursive_mutex g_api_lock;
bool g_cuda_ready = false;

[[noreturn]] void raise_last() { throw std::runtime_error(b200_last_error()); }

inline void check(int rc) {
    if (rc != B200_OK)
        raise_last();
}

inline void require_cuda(JitBackend backend, const char *what) {
    if (backend != JitBackend::CUDA)
        throw std::runtime_error(std::string(what) +
                                 "(): this build only provides the CUDA backend "
                                 "(no CPU fallback).");
}

} // namespace

void jit_init(uint32_t backends) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    // Like jitc_init (src/init.cpp:67-114) a backend that fails to initialise
    // is left disabled rather than raising; jit_has_backend reports it.
    if (backends & (1u << (uint32_t) JitBackend::CUDA))
        g_cuda_ready = b200_init() == B200_OK && b200_stream() != nullptr;
}

int jit_has_backend(JitBackend backend) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    return backend == JitBackend::CUDA && g_cuda_ready;
}

void jit_shutdown(int) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    b200_shutdown();
    g_cuda_ready = false;
}

void jit_sync_thread() {
    void *stream;
    {
        std::lock_guard<std::recursive_mutex> guard(g_api_lock);
        stream = b200_stream();
    }
    if (stream)
        check(b200_sync(stream)); // lock not held while blocking (src/init.cpp:516-517)
}

void jit_cuda_sync_stream(uintptr_t stream) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    check(b200_sync_stream((void *) stream));
}

void jit_set_flags(uint32_t flags) { b200_set_flags(flags); }
uint32_t jit_flags() { return b200_flags(); }
void jit_set_flag(JitFlag flag, int enable) { b200_set_flag((uint32_t) flag, enable); }
int jit_flag(JitFlag flag) { return (b200_flags() & (uint32_t) flag) ? 1 : 0; }

KernelHistoryEntry *jit_kernel_history() {
    // src/init.cpp jitc_kernel_history: waits for the events, returns a malloc'ed array
    // terminated by an entry whose backend is None; NULL when the history is empty
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    std::vector<B200KernelRecord> recs(64);
    int n = b200_kernel_history(recs.data(), (int) recs.size());
    // (records beyond the capacity are dropped by the C-ABI call; ask generously)
    if (n > (int) recs.size())
        n = (int) recs.size();
    if (n == 0)
        return nullptr;
    KernelHistoryEntry *out = (KernelHistoryEntry *) calloc((size_t) n + 1, sizeof(KernelHistoryEntry));
    for (int i = 0; i < n; ++i) {
        out[i].backend = JitBackend::CUDA;
        out[i].type = (KernelType) recs[i].type;
        out[i].recording_mode = KernelRecordingMode::Inactive;
        out[i].size = (uint32_t) recs[i].size;
        out[i].input_count = 1;
        out[i].output_count = 1;
        out[i].execution_time = recs[i].execution_time_ms;
    }
    return out;
}

void jit_kernel_history_clear() {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    b200_kernel_history_clear();
}

void *jit_malloc_migrate(void *ptr, JitBackend backend, int move) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    if (!ptr)
        return nullptr;
    int kind;
    if (backend == JitBackend::CUDA)
        kind = 0;
    else if (backend == JitBackend::None)
        kind = 1; // host: pinned memory
    else
        throw std::runtime_error("jit_malloc_migrate(): this build only provides the CUDA backend "
                                 "(no CPU fallback).");
    void *r = b200_malloc_migrate(ptr, kind, move);
    if (!r)
        raise_last();
    return r;
}

int jit_cuda_device_count() {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    return b200_device_count();
}

void jit_cuda_set_device(int device) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    check(b200_set_device(device));
}

int jit_cuda_device() {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    return b200_device();
}

void *jit_cuda_stream() {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    return b200_stream();
}

void *jit_malloc(JitBackend backend, size_t size, int shared) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_malloc");
    if (size == 0)
        return nullptr;
    void *ptr = b200_malloc(size, shared ? 1 : 0);
    if (!ptr)
        raise_last();
    return ptr;
}

void jit_free(void *ptr) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    check(b200_free(ptr));
}

void jit_memcpy(JitBackend backend, void *dst, const void *src, size_t size) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_memcpy");
    check(b200_memcpy(dst, src, size));
}

void jit_memcpy_async(JitBackend backend, void *dst, const void *src, size_t size) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_memcpy_async");
    check(b200_memcpy_async(nullptr, dst, src, size));
}

void jit_memset_async(JitBackend backend, void *ptr, uint32_t size, uint32_t isize,
                      const void *src) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_memset_async");
    check(b200_memset_async(nullptr, ptr, size, isize, src));
}

uint64_t jit_reduce_identity(VarType vt, ReduceOp op) {
    return b200_reduce_identity((int) vt, (int) op);
}

int jit_can_scatter_reduce(JitBackend backend, VarType vt, ReduceOp op) {
    require_cuda(backend, "jit_can_scatter_reduce");
    return b200_can_scatter_reduce((int) vt, (int) op);
}

void jit_reduce(JitBackend backend, VarType type, ReduceOp op, const void *in, uint32_t size,
                void *out) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_reduce");
    check(b200_reduce(nullptr, (int) type, (int) op, in, size, out));
}

void jit_reduce(JitBackend backend, VarType type, ReduceOp op, uint32_t size, const void *in,
                void *out) {
    jit_reduce(backend, type, op, in, size, out);
}

void jit_block_reduce(JitBackend backend, VarType type, ReduceOp op, uint32_t size,
                      uint32_t block_size, const void *in, void *out) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_block_reduce");
    check(b200_block_reduce(nullptr, (int) type, (int) op, size, block_size, in, out));
}

void jit_block_prefix_reduce(JitBackend backend, VarType type, ReduceOp op, uint32_t size,
                             uint32_t block_size, int exclusive, int reverse, const void *in,
                             void *out) {
    std::lock_guard<std::recursive_mutex> guard(g_api_lock);
    require_cuda(backend, "jit_block_prefix_reduce");
    check(b200_block_prefix_reduce(nullptr, (int) type, (int) op, size, block_size, exclusive,
                                   reverse, in, out));
}

// The two synchronising primitives hold the API lock for the whole enqueue (temporary
// allocations, per-device caches, launches) and drop it only while waiting for the
// stream -- state.lock + unlock_guard of the reference (src/cuda_ts.cpp:759, :964-967).

uint32_t jit_compress(JitBackend backend, const uint8_t *in, uint32_t size, uint32_t *out) {
    require_cuda(backend, "jit_compress");
    if (size == 0)
        return 0;
    if (b200_flags() & B200_FLAG_FORBID_SYNCHRONIZATION)
        throw std::runtime_error("Attempted to synchronize in a context, where synchronization "
                                 "was explicitly forbidden!");
    // the count lands in pinned memory owned by this call
    uint32_t *count_pinned = (uint32_t *) jit_malloc(JitBackend::CUDA, sizeof(uint32_t), 1);
    void *stream;
    int rc;
    {
        std::lock_guard<std::recursive_mutex> guard(g_api_lock);
        stream = b200_stream();
        rc = b200_compress_async(stream, in, size, out, count_pinned);
    }
    if (rc == B200_OK)
        rc = b200_sync(stream); // lock not held while blocking
    uint32_t count = rc == B200_OK ? *count_pinned : 0;
    jit_free(count_pinned);
    check(rc);
    return count;
}

uint32_t jit_block_mkperm(JitBackend backend, const uint32_t *values, uint32_t size,
                          uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                          uint32_t *offsets) {
    require_cuda(backend, "jit_block_mkperm");
    if (size == 0)
        return 0;
    // (no ForbidSynchronization check: the reference waits on an event here without
    // one, src/cuda_ts.cpp:964-967)
    const bool waits = offsets && (block_size >= size);
    void *stream;
    {
        std::lock_guard<std::recursive_mutex> guard(g_api_lock);
        stream = b200_stream();
        check(b200_block_mkperm_async(stream, values, size, block_size, bucket_count, perm, offsets));
    }
    if (!waits)
        return 0;
    if (cudaStreamSynchronize((cudaStream_t) stream) != cudaSuccess) // lock not held while blocking
        throw std::runtime_error("jit_block_mkperm(): stream synchronisation failed");
    return offsets[4 * (size_t) bucket_count];
}
