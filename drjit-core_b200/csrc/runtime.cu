// runtime.cu -- minimal runtime around the primitive kernels: device / stream
// bookkeeping, error reporting, allocation, copies and typed fills.
//
// Replaces, for this path only: jitc_cuda_init (src/cuda_core.cpp:266-539, no
// blob decompression / lazy PTX JIT any more -- the sm_100a SASS is linked into
// this library), jitc_malloc's rounding contract (src/malloc.cpp:113-124),
// CUDAThreadState::memset_async / memcpy(_async) (src/cuda_ts.cpp:129-183,
// :977-990) and the fill_64 kernel (resources/misc.cuh:28-32).
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace b200 {

static thread_local std::string tls_error;

struct DeviceState {
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
};

static std::mutex g_lock;
static std::atomic<bool> g_initialised{ false };
// Sized once by b200_init() and only cleared by b200_shutdown() (under g_lock);
// DeviceState::stream is published with release / acquire semantics.
static std::vector<DeviceState> g_devices;
static std::atomic<uint64_t> g_launches{ 0 };
static std::atomic<uint32_t> g_flags{ 0 };
// the stream the calling thread's current API call runs on (JitFlag::LaunchBlocking)
static thread_local cudaStream_t tls_stream = nullptr;
// pointer -> where it came from, for b200_free
struct AllocInfo {
    int kind;            // 0 device, 1 pinned host
    int device;          // device that was current at allocation time
    cudaStream_t stream; // stream the allocation was ordered on
    size_t bytes;        // rounded size
};
static std::unordered_map<void *, AllocInfo> g_allocs;

// ---- kernel history (JitFlag::KernelHistory, jit.h:2597-2709; src/cuda_ts.cpp:23-46)
struct HistoryRecord {
    int type;
    uint64_t size;
    cudaEvent_t start, end;
};
static std::vector<HistoryRecord> g_history; // g_lock

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    tls_error = buf;
    return code;
}

int cuda_fail(cudaError_t err, const char *what) {
    if (err == cudaSuccess)
        return B200_OK;
    return fail(B200_ERR_CUDA, "CUDA error %d (%s) in %s", (int) err,
                cudaGetErrorString(err), what);
}

void count_launch(uint64_t n) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    // JitFlag::LaunchBlocking: "force synchronization after every kernel launch"
    // (jit.h:1736-1739; src/cuda_ts.cpp:37-38)
    if (g_flags.load(std::memory_order_relaxed) & B200_FLAG_LAUNCH_BLOCKING)
        cudaStreamSynchronize(tls_stream);
}

uint32_t flags() { return g_flags.load(std::memory_order_relaxed); }

int sync_forbidden() {
    // src/init.cpp:503-505
    if (g_flags.load(std::memory_order_relaxed) & B200_FLAG_FORBID_SYNCHRONIZATION)
        return fail(B200_ERR_SYNC_FORBIDDEN,
                    "Attempted to synchronize in a context, where synchronization was "
                    "explicitly forbidden!");
    return B200_OK;
}

HistoryScope::HistoryScope(cudaStream_t stream_, int type_, uint64_t size_)
    : stream(stream_), type(type_), size(size_), start(nullptr) {
    if (!(g_flags.load(std::memory_order_relaxed) & B200_FLAG_KERNEL_HISTORY))
        return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess || cudaEventRecord(ev, stream) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    start = ev;
}

HistoryScope::~HistoryScope() {
    if (!start)
        return;
    cudaEvent_t end;
    if (cudaEventCreate(&end) != cudaSuccess || cudaEventRecord(end, stream) != cudaSuccess) {
        cudaGetLastError();
        cudaEventDestroy((cudaEvent_t) start);
        return;
    }
    std::lock_guard<std::mutex> guard(g_lock);
    g_history.push_back({ type, size, (cudaEvent_t) start, end });
}

const char *type_name(int vt) {
    // src/var.cpp type_name table
    static const char *names[16] = { "void", "bool", "?", "int8", "uint8", "int16",
        "uint16", "int32", "uint32", "int64", "uint64", "pointer", "?", "float16",
        "float32", "float64" };
    return (vt >= 0 && vt < 16) ? names[vt] : "?";
}

const char *op_name(int op) {
    static const char *names[7] = { "none", "add", "mul", "min", "max", "and", "or" };
    return (op >= 0 && op < 7) ? names[op] : "?";
}

// The CUDA runtime's current device of the calling thread is the single source
// of truth (b200_set_device == cudaSetDevice + lazy per-device setup), so the
// library follows whatever device the host program -- e.g. torch -- selected.
static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return dev;
}

int ensure_init() {
    if (g_initialised.load(std::memory_order_acquire))
        return B200_OK;
    return b200_init();
}

static int prepare_device(int dev) {
    std::lock_guard<std::mutex> guard(g_lock);
    if (dev < 0 || dev >= (int) g_devices.size())
        return fail(B200_ERR_INVALID, "b200: invalid device %d (library shut down?)", dev);
    DeviceState &d = g_devices[dev];
    if (d.stream)
        return B200_OK;
    if (d.cc_major < 10)
        return fail(B200_ERR_CUDA,
                    "b200: device %d has compute capability %d.%d; the kernels in "
                    "this library are built for sm_100a only",
                    dev, d.cc_major, d.cc_minor);
    B200_CUDA_CHECK(cudaSetDevice(dev));
    // A BLOCKING stream, like the reference's (cuStreamCreate(.., CU_STREAM_DEFAULT),
    // src/cuda_core.cpp:480): work on it is implicitly ordered with the legacy
    // default stream, so callers that stage inputs with cudaMemset / cudaMemcpy / their
    // own stream-0 kernels, or read results with cudaMemcpy right after a primitive,
    // keep the ordering they get from the reference ("Dr.Jit implicitly synchronizes
    // with respect to [the NULL stream]", jit.h:249-250).
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&d.stream, cudaStreamDefault));
    return B200_OK;
}

static cudaStream_t library_stream() {
    int dev = current_device();
    if (prepare_device(dev) != B200_OK)
        return nullptr;
    std::lock_guard<std::mutex> guard(g_lock);
    return dev < (int) g_devices.size() ? g_devices[dev].stream : nullptr;
}

cudaStream_t resolve_stream(void *stream) {
    cudaStream_t s = stream ? (cudaStream_t) stream : library_stream();
    tls_stream = s;
    return s;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = current_device();
    if (dev != cached_dev) {
        std::lock_guard<std::mutex> guard(g_lock);
        if (dev >= 0 && dev < (int) g_devices.size()) {
            cached = g_devices[dev].sm_count;
            cached_dev = dev;
        }
    }
    return cached;
}

// Freed temporaries must stay cached in the device's default pool: with the
// default release threshold (0) every synchronisation hands them back to the
// driver and the next cudaMallocAsync costs milliseconds.  Done once per device,
// for whichever device / stream the caller uses (jitc_malloc caches as well).
static std::atomic<bool> g_pool_ready[64];

static void retain_pool(int dev) {
    if (dev < 0 || dev >= 64 || g_pool_ready[dev].load(std::memory_order_acquire))
        return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    } else {
        cudaGetLastError();
    }
    g_pool_ready[dev].store(true, std::memory_order_release);
}

void *temp_alloc(size_t bytes, cudaStream_t stream) {
    void *ptr = nullptr;
    if (bytes == 0)
        bytes = 16;
    retain_pool(current_device());
    if (cudaMallocAsync(&ptr, bytes, stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return ptr;
}

void temp_free(void *ptr, cudaStream_t stream) {
    if (ptr)
        cudaFreeAsync(ptr, stream);
}

// Two ticket words per (device, stream), handed out from a per-device block of device
// memory that is zeroed once.
unsigned int *stream_ticket(cudaStream_t stream) {
    constexpr int SLOTS = 1024;
    struct Table {
        unsigned int *words = nullptr;
        std::vector<cudaStream_t> owners;
    };
    static std::mutex mutex;
    static std::vector<Table> tables;
    const int dev = current_device();
    std::lock_guard<std::mutex> guard(mutex);
    if ((int) tables.size() <= dev)
        tables.resize(dev + 1);
    Table &t = tables[dev];
    if (!t.words) {
        if (cudaMalloc((void **) &t.words, SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(t.words, 0, SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess) {
            cudaGetLastError();
            t.words = nullptr;
            return nullptr;
        }
    }
    for (size_t i = 0; i < t.owners.size(); ++i)
        if (t.owners[i] == stream)
            return t.words + 2 * i;
    if ((int) t.owners.size() == SLOTS)
        return nullptr;
    t.owners.push_back(stream);
    return t.words + 2 * (t.owners.size() - 1);
}

// One pinned host word per calling thread.  Kernels of the synchronising entry
// points (jit_compress) write their scalar result straight into it (pinned memory
// is device-accessible under unified addressing), so the host reads it after the
// stream synchronisation without a device-to-host copy into pageable memory.  The
// reference reads its count from pinned memory in the same way
// (src/cuda_ts.cpp:759-762).
uint32_t *pinned_scalar() {
    static thread_local uint32_t *word = nullptr;
    if (!word) {
        if (cudaHostAlloc((void **) &word, 64, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            word = nullptr;
        }
    }
    return word;
}

// Typed fill: 'count' elements of 2 / 4 / 8 bytes.  16-byte stores in the body,
// element stores for the unaligned head and the tail.
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T *ptr, uint64_t count, T value) {
    constexpr uint32_t N = 16 / sizeof(T);
    uint64_t head = ((16 - ((uintptr_t) ptr & 15)) & 15) / sizeof(T);
    if (head > count)
        head = count;
    uint64_t nvec = (count - head) / N;
    uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x,
             stride = (uint64_t) gridDim.x * blockDim.x;
    Vec16<T> v;
    #pragma unroll
    for (uint32_t i = 0; i < N; ++i)
        v.elem[i] = value;
    uint4 *body = (uint4 *) (ptr + head);
    for (uint64_t i = tid; i < nvec; i += stride)
        st_stream(body + i, v.raw);
    uint64_t tail_start = head + nvec * N;
    for (uint64_t i = tid; i < head; i += stride)
        ptr[i] = value;
    for (uint64_t i = tail_start + tid; i < count; i += stride)
        ptr[i] = value;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_init(void) {
    std::lock_guard<std::mutex> guard(g_lock);
    if (g_initialised.load(std::memory_order_acquire))
        return B200_OK;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(B200_ERR_CUDA,
                    "b200_init(): no CUDA device available (%s); this library has "
                    "no CPU fallback",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    }
    int prev = 0;
    cudaGetDevice(&prev);
    g_devices.resize(count);
    for (int i = 0; i < count; ++i) {
        cudaDeviceProp prop;
        B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, i));
        g_devices[i].sm_count = prop.multiProcessorCount;
        g_devices[i].cc_major = prop.major;
        g_devices[i].cc_minor = prop.minor;
        // Streams / pools are created lazily per device on first selection so
        // that a rank of a multi-process job only ever touches its own GPU.
    }
    cudaSetDevice(prev);
    g_initialised.store(true, std::memory_order_release);
    return B200_OK;
}


static void release_pinned_cache(); // (allocator, below)

int b200_shutdown(void) {
    std::lock_guard<std::mutex> guard(g_lock);
    if (!g_initialised.load(std::memory_order_acquire))
        return B200_OK;
    release_pinned_cache();
    for (HistoryRecord &h : g_history) {
        cudaEventDestroy(h.start);
        cudaEventDestroy(h.end);
    }
    g_history.clear();
    for (size_t i = 0; i < g_devices.size(); ++i) {
        if (g_devices[i].stream) {
            cudaSetDevice((int) i);
            cudaStreamSynchronize(g_devices[i].stream);
            cudaStreamDestroy(g_devices[i].stream);
            g_devices[i].stream = nullptr;
        }
    }
    g_devices.clear();
    g_initialised.store(false, std::memory_order_release);
    return B200_OK;
}

const char *b200_last_error(void) { return tls_error.c_str(); }

int b200_device_count(void) {
    if (ensure_init())
        return 0;
    std::lock_guard<std::mutex> guard(g_lock);
    return (int) g_devices.size();
}

int b200_set_device(int device) {
    int rc = ensure_init();
    if (rc)
        return rc;
    int count = b200_device_count();
    if (device < 0 || device >= count)
        return fail(B200_ERR_INVALID, "b200_set_device(%d): must be in the range 0..%d!",
                    device, count - 1);
    B200_CUDA_CHECK(cudaSetDevice(device));
    return prepare_device(device);
}

int b200_device(void) {
    if (ensure_init())
        return -1;
    return current_device();
}

void *b200_stream(void) {
    if (ensure_init())
        return nullptr;
    return (void *) library_stream();
}

int b200_sm_count(void) {
    if (ensure_init())
        return 0;
    return sm_count();
}

int b200_sync(void *stream) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if ((rc = sync_forbidden()))
        return rc;
    cudaStream_t s = resolve_stream(stream);
    B200_CUDA_CHECK(cudaStreamSynchronize(s));
    return B200_OK;
}

/* jit_cuda_sync_stream (jit.h:243-255; src/init.cpp): an event recorded on the
 * library stream, 'other' waits for it.  2 = the caller's per-thread default stream. */
int b200_sync_stream(void *other) {
    int rc = ensure_init();
    if (rc)
        return rc;
    cudaStream_t lib = library_stream();
    cudaEvent_t ev;
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t err = cudaEventRecord(ev, lib);
    if (err == cudaSuccess)
        err = cudaStreamWaitEvent((cudaStream_t) other, ev, 0);
    cudaEventDestroy(ev);
    return cuda_fail(err, "jit_cuda_sync_stream()");
}

void b200_set_flags(uint32_t flags) { g_flags.store(flags, std::memory_order_relaxed); }
uint32_t b200_flags(void) { return g_flags.load(std::memory_order_relaxed); }

void b200_set_flag(uint32_t flag, int enable) {
    if (enable)
        g_flags.fetch_or(flag, std::memory_order_relaxed);
    else
        g_flags.fetch_and(~flag, std::memory_order_relaxed);
}

int b200_kernel_history(B200KernelRecord *out, int capacity) {
    // like jit_kernel_history(): waits for the recorded events, hands the entries
    // over and clears the history (src/init.cpp jitc_kernel_history)
    std::vector<HistoryRecord> recs;
    {
        std::lock_guard<std::mutex> guard(g_lock);
        recs.swap(g_history);
    }
    int n = 0;
    for (HistoryRecord &h : recs) {
        float ms = 0.f;
        cudaEventSynchronize(h.end);
        cudaEventElapsedTime(&ms, h.start, h.end);
        cudaEventDestroy(h.start);
        cudaEventDestroy(h.end);
        if (out && n < capacity) {
            out[n].type = h.type;
            out[n].size = h.size;
            out[n].execution_time_ms = ms;
        }
        n++;
    }
    cudaGetLastError();
    return n;
}

void b200_kernel_history_clear(void) { b200_kernel_history(nullptr, 0); }


// Allocator contract of the reference (src/malloc.cpp:102-306):
// sizes are rounded up to a power of two >= 64 bytes (compress and all / any of the
// reference rely on that padding, jit.h:2382-2383), freed blocks are cached, and
// allocation / release are ordered on the device's stream instead of synchronising
// the device.  Device memory: the CUDA stream-ordered pool of the library stream
// (release threshold raised in retain_pool(), so freed blocks stay cached).  Pinned
// host memory: a free list per size; a block is handed out again once the event
// recorded at its release has completed.
struct PinnedBlock {
    void *ptr;
    cudaEvent_t released;
};
static std::unordered_map<size_t, std::vector<PinnedBlock>> g_pinned_cache;
// pointer -> rounded size (pinned blocks only)
static std::unordered_map<void *, size_t> g_pinned_size;

static void release_pinned_cache() { // g_lock held
    for (auto &kv : g_pinned_cache) {
        for (PinnedBlock &b : kv.second) {
            cudaEventSynchronize(b.released);
            cudaEventDestroy(b.released);
            cudaFreeHost(b.ptr);
        }
    }
    g_pinned_cache.clear();
}

void *b200_malloc(size_t size, int kind) {
    if (ensure_init() || size == 0)
        return nullptr;
    // src/malloc.cpp:113-124: round up to 64 bytes, then to a power of two
    size = (size + 63) / 64 * 64;
    size_t rounded = 64;
    while (rounded < size)
        rounded <<= 1;
    void *ptr = nullptr;
    const int dev = current_device();
    if (kind == 1) {
        {
            std::lock_guard<std::mutex> guard(g_lock);
            auto &list = g_pinned_cache[rounded];
            for (size_t i = 0; i < list.size(); ++i) {
                if (cudaEventQuery(list[i].released) == cudaSuccess) {
                    ptr = list[i].ptr;
                    cudaEventDestroy(list[i].released);
                    list.erase(list.begin() + i);
                    break;
                }
            }
            cudaGetLastError(); // cudaErrorNotReady from the queries
        }
        if (!ptr) {
            cudaError_t err = cudaMallocHost(&ptr, rounded);
            if (err != cudaSuccess) {
                cuda_fail(err, "cudaMallocHost");
                return nullptr;
            }
        }
        std::lock_guard<std::mutex> guard(g_lock);
        g_allocs[ptr] = { kind, dev, nullptr, rounded };
        g_pinned_size[ptr] = rounded;
        return ptr;
    }
    cudaStream_t s = library_stream();
    retain_pool(dev);
    cudaError_t err = cudaMallocAsync(&ptr, rounded, s);
    if (err != cudaSuccess) {
        cuda_fail(err, "cudaMallocAsync");
        return nullptr;
    }
    std::lock_guard<std::mutex> guard(g_lock);
    g_allocs[ptr] = { kind, dev, s, rounded };
    return ptr;
}

/// Release 'ptr' in the order of 'stream' (NULL: the library stream of the device
/// the block was allocated on).  Work that uses the block on ANOTHER stream must
/// be ordered before the release by the caller -- b200_free_on(that stream, ptr).
static int free_impl(void *ptr, bool have_stream, cudaStream_t user_stream) {
    if (!ptr)
        return B200_OK;
    AllocInfo info{};
    size_t pinned_size = 0;
    {
        std::lock_guard<std::mutex> guard(g_lock);
        auto it = g_allocs.find(ptr);
        if (it == g_allocs.end())
            return fail(B200_ERR_INVALID, "b200_free(): unknown address %p!", ptr);
        info = it->second;
        g_allocs.erase(it);
        if (info.kind == 1) {
            pinned_size = g_pinned_size[ptr];
            g_pinned_size.erase(ptr);
        }
    }
    // the block belongs to the device (and pool) it was allocated on, whatever device
    // is current now
    const int cur = current_device();
    if (cur != info.device)
        cudaSetDevice(info.device);
    cudaStream_t s = have_stream ? user_stream : (info.stream ? info.stream : library_stream());
    int rc = B200_OK;
    if (info.kind == 1) {
        // reusable once the work enqueued so far (which may still read / write it) is done
        cudaEvent_t ev;
        cudaError_t err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (err == cudaSuccess)
            err = cudaEventRecord(ev, s);
        if (err == cudaSuccess) {
            std::lock_guard<std::mutex> guard(g_lock);
            g_pinned_cache[pinned_size].push_back({ ptr, ev });
        } else {
            rc = cuda_fail(err, "b200_free(): event");
        }
    } else {
        rc = cuda_fail(cudaFreeAsync(ptr, s), "cudaFreeAsync");
    }
    if (cur != info.device)
        cudaSetDevice(cur);
    return rc;
}

int b200_free(void *ptr) { return free_impl(ptr, false, nullptr); }

int b200_free_on(void *stream, void *ptr) {
    return free_impl(ptr, stream != nullptr, (cudaStream_t) stream);
}

/* jit_malloc_migrate (jit.h:495-518) for the allocation kinds of this library:
 * kind 0 = device, 1 = pinned host.  Same kind + move: returned unchanged.
 * Otherwise a new block, an asynchronous copy on the library stream and (move)
 * the release of the source, also in stream order. */
void *b200_malloc_migrate(void *ptr, int kind, int move) {
    if (!ptr || ensure_init())
        return nullptr;
    AllocInfo info{};
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> guard(g_lock);
        auto it = g_allocs.find(ptr);
        if (it == g_allocs.end()) {
            fail(B200_ERR_INVALID, "jit_malloc_migrate(): unknown address %p!", ptr);
            return nullptr;
        }
        info = it->second;
        bytes = info.bytes;
    }
    if (info.kind == kind && (kind == 1 || info.device == current_device()) && move)
        return ptr;
    void *dst = b200_malloc(bytes, kind);
    if (!dst)
        return nullptr;
    cudaStream_t s = library_stream();
    if (cudaMemcpyAsync(dst, ptr, bytes, cudaMemcpyDefault, s) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "jit_malloc_migrate(): copy");
        b200_free(dst);
        return nullptr;
    }
    if (move)
        b200_free(ptr);
    return dst;
}

int b200_memcpy(void *dst, const void *src, size_t size) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if ((rc = sync_forbidden()))
        return rc;
    // synchronous with respect to the library stream as well (jit_memcpy syncs)
    cudaStream_t s = resolve_stream(nullptr);
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, size, cudaMemcpyDefault, s));
    B200_CUDA_CHECK(cudaStreamSynchronize(s));
    return B200_OK;
}

int b200_memcpy_async(void *stream, void *dst, const void *src, size_t size) {
    int rc = ensure_init();
    if (rc)
        return rc;
    cudaStream_t s = resolve_stream(stream);
    HistoryScope hs(s, B200_KERNEL_MEMCPY, size);
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, size, cudaMemcpyDefault, s));
    return B200_OK;
}

int b200_memset_async(void *stream, void *ptr, uint64_t size, uint32_t isize,
                      const void *src) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (isize != 1 && isize != 2 && isize != 4 && isize != 8)
        return fail(B200_ERR_INVALID,
                    "jit_memset_async(): invalid element size (must be 1, 2, 4, or 8)!");
    if (size == 0)
        return B200_OK;
    cudaStream_t s = resolve_stream(stream);
    HistoryScope hs(s, B200_KERNEL_MEMSET, size);

    // Patterns whose bytes are all equal collapse to a byte memset
    // (src/cuda_ts.cpp:143-148 does this for zero only)
    const uint8_t *b = (const uint8_t *) src;
    bool uniform = true;
    for (uint32_t i = 1; i < isize; ++i)
        uniform &= b[i] == b[0];
    if (uniform) {
        B200_CUDA_CHECK(cudaMemsetAsync(ptr, b[0], size * isize, s));
        return B200_OK;
    }

    if (((uintptr_t) ptr) % isize != 0)
        return fail(B200_ERR_INVALID, "jit_memset_async(): misaligned address %p!", ptr);

    uint64_t vecs = size * isize / 16 + 1;
    uint32_t blocks = (uint32_t) std::min<uint64_t>(ceil_div(vecs, 256), (uint64_t) sm_count() * 8);
    switch (isize) {
        case 2: { uint16_t v; memcpy(&v, src, 2); fill_kernel<uint16_t><<<blocks, 256, 0, s>>>((uint16_t *) ptr, size, v); break; }
        case 4: { uint32_t v; memcpy(&v, src, 4); fill_kernel<uint32_t><<<blocks, 256, 0, s>>>((uint32_t *) ptr, size, v); break; }
        case 8: { uint64_t v; memcpy(&v, src, 8); fill_kernel<uint64_t><<<blocks, 256, 0, s>>>((uint64_t *) ptr, size, v); break; }
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

uint64_t b200_reduce_identity(int vt, int op) {
    // src/var.cpp:140-166 and :2642-2652
    static const uint64_t all_ones[16] = {
        0, 1, 0, 0xff, 0xff, 0xffff, 0xffff, 0xffffffffu, 0xffffffffu,
        ~0ull, ~0ull, ~0ull, 0, 0xffff, 0xffffffffu, ~0ull };
    static const uint64_t one[16] = {
        0, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0x3c00, 0x3f800000,
        0x3ff0000000000000ull };
    static const uint64_t tmin[16] = {
        0, 0, 0, 0x80, 0, 0x8000, 0, 0x80000000u, 0, 0x8000000000000000ull, 0, 0,
        0, 0xfc00, 0xff800000u, 0xfff0000000000000ull };
    static const uint64_t tmax[16] = {
        0, 1, 0, 0x7f, 0xff, 0x7fff, 0xffff, 0x7fffffff, 0xffffffffu,
        0x7fffffffffffffffull, ~0ull, ~0ull, 0, 0x7c00, 0x7f800000,
        0x7ff0000000000000ull };
    if (vt < 0 || vt >= 16)
        return 0;
    switch (op) {
        case B200_OP_OR:
        case B200_OP_ADD: return 0;
        case B200_OP_AND: return all_ones[vt];
        case B200_OP_MUL: return one[vt];
        case B200_OP_MIN: return tmax[vt];
        case B200_OP_MAX: return tmin[vt];
        default: return 0;
    }
}

uint64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

} // extern "C"
