// common.cuh -- shared device/host helpers for the sm_100a primitive kernels.
//
// Replaces resources/common.h of the reference (functors :93-163, flag helpers
// :180-255).  Nothing here is specific to one primitive: element traits, the
// six reduction operators, 128-bit streaming loads/stores, warp collectives
// and the tile-descriptor protocol used by the single-pass scans.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/drjit_b200.h"

#define B200_DEVICE __device__ __forceinline__
#define FULL_MASK 0xffffffffu

// ---------------------------------------------------------------- host side

namespace b200 {

/// Set the thread-local error message and return 'code'
int fail(int code, const char *fmt, ...);
/// Record a CUDA error (returns B200_ERR_CUDA) or pass through success
int cuda_fail(cudaError_t err, const char *what);
/// Resolve the stream argument of an API call (NULL -> library stream)
cudaStream_t resolve_stream(void *stream);
/// Number of SMs of the current device
int sm_count();
/// Count kernel launches (telemetry for bench.py)
void count_launch(uint64_t n = 1);
/// Make sure the runtime is initialised (lazily calls b200_init)
int ensure_init();
/// Current JitFlag bits (b200_set_flags)
uint32_t flags();
/// B200_ERR_SYNC_FORBIDDEN (with the reference's message) when
/// JitFlag::ForbidSynchronization is set, else B200_OK.  Called by every entry
/// point that blocks on the stream (src/init.cpp:503-505).
int sync_forbidden();

/// JitFlag::KernelHistory: brackets one primitive call with two events on its
/// stream and appends a record of the given KernelType (jit.h:2597-2634) -- what
/// CUDAThreadState::submit does around every precompiled kernel of the reference
/// (src/cuda_ts.cpp:23-46).  A no-op when the flag is clear.
struct HistoryScope {
    HistoryScope(cudaStream_t stream, int type, uint64_t size);
    ~HistoryScope();
    HistoryScope(const HistoryScope &) = delete;
    HistoryScope &operator=(const HistoryScope &) = delete;
    cudaStream_t stream;
    int type;
    uint64_t size;
    void *start;
};

/// Stream-ordered temporary memory (CUDA memory-pool backed)
void *temp_alloc(size_t bytes, cudaStream_t stream);
void temp_free(void *ptr, cudaStream_t stream);
/// A pinned (device-accessible) host word owned by the calling thread, or NULL
uint32_t *pinned_scalar();
/// TWO zero-initialised device words that belong to (current device, stream): kernels
/// on that stream use them as ticket / "last CTA done" counters and leave them at zero
/// again, so consecutive launches (in-order on the stream, also when replayed from a
/// graph captured on it) need no memset.  NULL when the table is full.
unsigned int *stream_ticket(cudaStream_t stream);

#define B200_CUDA_CHECK(expr)                                                  \
    do {                                                                       \
        cudaError_t err__ = (expr);                                            \
        if (err__ != cudaSuccess)                                              \
            return b200::cuda_fail(err__, #expr);                              \
    } while (0)

#define B200_LAUNCH_CHECK()                                                    \
    do {                                                                       \
        b200::count_launch();                                                  \
        cudaError_t err__ = cudaGetLastError();                                \
        if (err__ != cudaSuccess)                                              \
            return b200::cuda_fail(err__, "kernel launch");                    \
    } while (0)

inline uint32_t type_size(int vt) {
    // src/var.cpp:117-119
    static const uint32_t ts[16] = { 0, 1, 0, 1, 1, 2, 2, 4, 4, 8, 8, 8, 0, 2, 4, 8 };
    return (vt >= 0 && vt < 16) ? ts[vt] : 0;
}

const char *type_name(int vt);
const char *op_name(int op);

inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }
inline uint32_t log2i(uint64_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }

} // namespace b200

// -------------------------------------------------------------- device side
#if defined(__CUDACC__)

namespace b200 {

// Reduction operators --------------------------------------------------------
//
// Value is the type arithmetic is carried out in: half accumulates in float
// (reference: resources/common.h:93-107), everything else in its own type.

template <typename T> struct ValueOf { using type = T; };
template <> struct ValueOf<__half> { using type = float; };

template <typename T> B200_DEVICE typename ValueOf<T>::type to_value(T x) { return x; }
template <> B200_DEVICE float to_value<__half>(__half x) { return __half2float(x); }

template <typename T> B200_DEVICE T from_value(typename ValueOf<T>::type v) { return v; }
template <> B200_DEVICE __half from_value<__half>(float v) { return __float2half_rn(v); }

template <typename V> struct Limits;
template <> struct Limits<uint8_t>  { static B200_DEVICE uint8_t lo() { return 0; } static B200_DEVICE uint8_t hi() { return 0xff; } };
template <> struct Limits<uint32_t> { static B200_DEVICE uint32_t lo() { return 0; } static B200_DEVICE uint32_t hi() { return 0xffffffffu; } };
template <> struct Limits<int32_t>  { static B200_DEVICE int32_t lo() { return (int32_t) 0x80000000; } static B200_DEVICE int32_t hi() { return 0x7fffffff; } };
template <> struct Limits<uint64_t> { static B200_DEVICE uint64_t lo() { return 0; } static B200_DEVICE uint64_t hi() { return ~0ull; } };
template <> struct Limits<int64_t>  { static B200_DEVICE int64_t lo() { return (int64_t) 0x8000000000000000ull; } static B200_DEVICE int64_t hi() { return 0x7fffffffffffffffll; } };
template <> struct Limits<unsigned long long> { static B200_DEVICE unsigned long long lo() { return 0; } static B200_DEVICE unsigned long long hi() { return ~0ull; } };
template <> struct Limits<long long> { static B200_DEVICE long long lo() { return (long long) 0x8000000000000000ull; } static B200_DEVICE long long hi() { return 0x7fffffffffffffffll; } };
template <> struct Limits<float>    { static B200_DEVICE float lo() { return -__int_as_float(0x7f800000); } static B200_DEVICE float hi() { return __int_as_float(0x7f800000); } };
template <> struct Limits<double>   { static B200_DEVICE double lo() { return -__longlong_as_double(0x7ff0000000000000ll); } static B200_DEVICE double hi() { return __longlong_as_double(0x7ff0000000000000ll); } };

B200_DEVICE float  vmin(float a, float b) { return fminf(a, b); }
B200_DEVICE double vmin(double a, double b) { return fmin(a, b); }
B200_DEVICE float  vmax(float a, float b) { return fmaxf(a, b); }
B200_DEVICE double vmax(double a, double b) { return fmax(a, b); }
template <typename V> B200_DEVICE V vmin(V a, V b) { return a < b ? a : b; }
template <typename V> B200_DEVICE V vmax(V a, V b) { return a > b ? a : b; }

/// The six ReduceOp operators on a value type V (jit.h:990-1014)
template <typename V, int Op> struct Red {
    static constexpr bool is_int = std::is_integral<V>::value;

    static B200_DEVICE V identity() {
        if constexpr (Op == B200_OP_ADD || Op == B200_OP_OR) return (V) 0;
        else if constexpr (Op == B200_OP_MUL) return (V) 1;
        else if constexpr (Op == B200_OP_MIN) return Limits<V>::hi();
        else if constexpr (Op == B200_OP_MAX) return Limits<V>::lo();
        else /* AND */ {
            if constexpr (is_int) return (V) ~(V) 0; else return (V) 0;
        }
    }

    static B200_DEVICE V apply(V a, V b) {
        if constexpr (Op == B200_OP_ADD) return (V) (a + b);
        else if constexpr (Op == B200_OP_MUL) return (V) (a * b);
        else if constexpr (Op == B200_OP_MIN) return vmin(a, b);
        else if constexpr (Op == B200_OP_MAX) return vmax(a, b);
        else if constexpr (Op == B200_OP_AND) {
            if constexpr (is_int) return (V) (a & b); else return a;
        } else {
            if constexpr (is_int) return (V) (a | b); else return a;
        }
    }
};

// 128-bit streaming memory access --------------------------------------------

/// 16-byte global load that does not pollute L1 (data is touched once)
B200_DEVICE uint4 ld_stream(const void *ptr) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
    return r;
}

/// 16-byte global load, coherent (used when in == out is possible)
B200_DEVICE uint4 ld_stream_coherent(const void *ptr) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr) : "memory");
    return r;
}

/// 16-byte streaming global store
B200_DEVICE void st_stream(void *ptr, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

/// Number of elements of T in a 16-byte vector
template <typename T> struct VecInfo { static constexpr int N = 16 / sizeof(T); };

/// Reinterpret a 16-byte register vector as N elements of T
template <typename T> union Vec16 {
    uint4 raw;
    T elem[16 / sizeof(T)];
    B200_DEVICE Vec16() {}
};

// Warp collectives -------------------------------------------------------------

template <typename V> B200_DEVICE V shfl_xor(V v, int delta) {
    if constexpr (sizeof(V) == 8) {
        uint64_t u; memcpy(&u, &v, 8);
        uint32_t lo = (uint32_t) u, hi = (uint32_t) (u >> 32);
        lo = __shfl_xor_sync(FULL_MASK, lo, delta);
        hi = __shfl_xor_sync(FULL_MASK, hi, delta);
        u = ((uint64_t) hi << 32) | lo;
        V r; memcpy(&r, &u, 8); return r;
    } else if constexpr (sizeof(V) == 4) {
        uint32_t u; memcpy(&u, &v, 4);
        u = __shfl_xor_sync(FULL_MASK, u, delta);
        V r; memcpy(&r, &u, 4); return r;
    } else {
        uint32_t u = (uint32_t) v;
        u = __shfl_xor_sync(FULL_MASK, u, delta);
        return (V) u;
    }
}

template <typename V> B200_DEVICE V shfl_up(V v, int delta) {
    if constexpr (sizeof(V) == 8) {
        uint64_t u; memcpy(&u, &v, 8);
        uint32_t lo = (uint32_t) u, hi = (uint32_t) (u >> 32);
        lo = __shfl_up_sync(FULL_MASK, lo, delta);
        hi = __shfl_up_sync(FULL_MASK, hi, delta);
        u = ((uint64_t) hi << 32) | lo;
        V r; memcpy(&r, &u, 8); return r;
    } else if constexpr (sizeof(V) == 4) {
        uint32_t u; memcpy(&u, &v, 4);
        u = __shfl_up_sync(FULL_MASK, u, delta);
        V r; memcpy(&r, &u, 4); return r;
    } else {
        uint32_t u = (uint32_t) v;
        u = __shfl_up_sync(FULL_MASK, u, delta);
        return (V) u;
    }
}

template <typename V> B200_DEVICE V shfl_idx(V v, int lane) {
    if constexpr (sizeof(V) == 8) {
        uint64_t u; memcpy(&u, &v, 8);
        uint32_t lo = (uint32_t) u, hi = (uint32_t) (u >> 32);
        lo = __shfl_sync(FULL_MASK, lo, lane);
        hi = __shfl_sync(FULL_MASK, hi, lane);
        u = ((uint64_t) hi << 32) | lo;
        V r; memcpy(&r, &u, 8); return r;
    } else if constexpr (sizeof(V) == 4) {
        uint32_t u; memcpy(&u, &v, 4);
        u = __shfl_sync(FULL_MASK, u, lane);
        V r; memcpy(&r, &u, 4); return r;
    } else {
        uint32_t u = (uint32_t) v;
        u = __shfl_sync(FULL_MASK, u, lane);
        return (V) u;
    }
}

/// Butterfly all-reduce over the 'width' (power of two, <= 32) lanes that share
/// the same lane / width.  32-bit integer add/min/max/and/or map to redux.sync
/// when the whole warp participates.
template <typename V, int Op> B200_DEVICE V warp_reduce(V v, int width = 32) {
    if constexpr (std::is_integral<V>::value && sizeof(V) == 4 && Op != B200_OP_MUL) {
        if (width == 32) {
            if constexpr (Op == B200_OP_ADD) return (V) __reduce_add_sync(FULL_MASK, (uint32_t) v);
            else if constexpr (Op == B200_OP_AND) return (V) __reduce_and_sync(FULL_MASK, (uint32_t) v);
            else if constexpr (Op == B200_OP_OR) return (V) __reduce_or_sync(FULL_MASK, (uint32_t) v);
            else if constexpr (Op == B200_OP_MIN) return (V) __reduce_min_sync(FULL_MASK, v);
            else return (V) __reduce_max_sync(FULL_MASK, v);
        }
    }
    for (int d = width >> 1; d > 0; d >>= 1)
        v = Red<V, Op>::apply(v, shfl_xor(v, d));
    return v;
}

// Tile descriptors for single-pass (decoupled look-back) scans -----------------
//
// One 64-bit word carries {status, 32-bit payload} and is published / observed
// with a single relaxed gpu-scope access, so no fence is needed between value
// and flag.  64-bit payloads use one 16-byte {value, status} pair, published /
// observed with ONE 128-bit relaxed access (st / ld.relaxed.gpu.global.b128 ->
// STG / LDG.E.128.STRONG.GPU): a reader never sees a torn pair and does not have
// to re-poll until two tags agree.  (The reference uses ld.volatile / st.cg
// status+value pairs, resources/common.h:180-255.)

enum : uint32_t { DESC_INVALID = 0, DESC_AGGREGATE = 1, DESC_PREFIX = 2 };

B200_DEVICE void st_relaxed_u64(uint64_t *ptr, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(ptr), "l"(v) : "memory");
}

B200_DEVICE uint64_t ld_relaxed_u64(const uint64_t *ptr) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}

/// 16-byte aligned, single-copy atomic 128-bit store / load
B200_DEVICE void st_relaxed_b128(uint64_t *ptr, uint64_t lo, uint64_t hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}"
                 :: "l"(ptr), "l"(lo), "l"(hi) : "memory");
}

B200_DEVICE void ld_relaxed_b128(const uint64_t *ptr, uint64_t &lo, uint64_t &hi) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(lo), "=l"(hi) : "l"(ptr) : "memory");
}

template <typename V> struct Desc {
    static constexpr int WORDS = sizeof(V) == 8 ? 2 : 1;

    static B200_DEVICE void publish(uint64_t *base, uint32_t tile, uint32_t status, V value) {
        uint64_t *p = base + (size_t) tile * WORDS;
        if constexpr (WORDS == 1) {
            uint32_t bits = 0;
            memcpy(&bits, &value, sizeof(V));
            st_relaxed_u64(p, ((uint64_t) status << 32) | bits);
        } else {
            uint64_t bits;
            memcpy(&bits, &value, 8);
            st_relaxed_b128(p, bits, (uint64_t) status);
        }
    }

    /// Returns the status (DESC_INVALID while the entry is not yet consistent)
    static B200_DEVICE uint32_t observe(const uint64_t *base, uint32_t tile, V &value) {
        const uint64_t *p = base + (size_t) tile * WORDS;
        if constexpr (WORDS == 1) {
            uint64_t w = ld_relaxed_u64(p);
            uint32_t bits = (uint32_t) w;
            memcpy(&value, &bits, sizeof(V));
            return (uint32_t) (w >> 32);
        } else {
            uint64_t bits, st;
            ld_relaxed_b128(p, bits, st);
            memcpy(&value, &bits, 8);
            return (uint32_t) st;
        }
    }
};

} // namespace b200

#endif // __CUDACC__
