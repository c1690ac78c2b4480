// packet.cu -- the packet (array-of-structures) accesses next to the scatter path and
// the scatter variants that scatter.cu does not cover:
//
//   b200_scatter_packet     jit_var_scatter_packet WITHOUT reduction (emitter
//                           jitc_cuda_render_scatter_packet, src/cuda_packet.cpp:329-443):
//                           target[index[i] * W + k] = values[k][i], one vector store of
//                           up to 128 bits per 16 bytes of packet;
//   b200_gather_packet      jit_var_gather_packet (jitc_cuda_render_gather_packet,
//                           src/cuda_packet.cpp:18-166): out[k][i] = source[index[i] * W + k]
//                           with vector loads, masked-off lanes read 0;
//   f16 packet reductions   red.global.v2 / .v4 / .v8.f16.add.noftz (cuda_packet.cpp:229-266,
//                           available from compute capability 9.0);
//   b200_scatter_reduce_idx plain scatter (ReduceOp::Identity) and the index types the
//                           reference accepts besides uint32 (int32 / uint64 / int64,
//                           src/op.cpp:2899-3086).
//
// A packet is handled as W * sizeof(T) raw bytes: the W separate (structure-of-arrays)
// value streams are read / written coalesced, the packet itself moves in the widest
// naturally aligned chunks (16, 8, 4, 2 or 1 bytes) that both its size and the base
// address allow.  64-bit indices: scatter_reduce_wide in scatter.cu.
#include "common.cuh"

#include <algorithm>
#include <cuda_fp16.h>

namespace b200 {

static constexpr int PK_THREADS = 256;

template <int BYTES> struct Chunk;
template <> struct Chunk<16> { using type = uint4; };
template <> struct Chunk<8> { using type = uint2; };
template <> struct Chunk<4> { using type = uint32_t; };
template <> struct Chunk<2> { using type = uint16_t; };
template <> struct Chunk<1> { using type = uint8_t; };

template <int W> struct SoaPtrs { void *v[W]; };

template <int TS> struct Raw;
template <> struct Raw<1> { using type = uint8_t; };
template <> struct Raw<2> { using type = uint16_t; };
template <> struct Raw<4> { using type = uint32_t; };
template <> struct Raw<8> { using type = uint64_t; };

/// target[index[i] * W + k] = values[k][i]; CB: bytes per store
template <int TS, int W, int CB>
__global__ void __launch_bounds__(PK_THREADS)
scatter_packet_store_kernel(uint8_t *__restrict__ target, const SoaPtrs<W> values,
                            const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask, uint64_t n) {
    using R = typename Raw<TS>::type;
    using C = typename Chunk<CB>::type;
    constexpr int PB = TS * W;
    static_assert(PB % CB == 0, "whole chunks");
    const uint64_t stride = (uint64_t) gridDim.x * PK_THREADS;
    for (uint64_t i = (uint64_t) blockIdx.x * PK_THREADS + threadIdx.x; i < n; i += stride) {
        if (mask && __ldcs(mask + i) == 0)
            continue;
        union { R r[W]; C c[PB / CB]; } u;
        #pragma unroll
        for (int k = 0; k < W; ++k)
            u.r[k] = __ldcs((const R *) values.v[k] + i);
        C *p = (C *) (target + (uint64_t) __ldcs(index + i) * PB);
        #pragma unroll
        for (int q = 0; q < PB / CB; ++q)
            p[q] = u.c[q];
    }
}

/// out[k][i] = mask[i] ? source[index[i] * W + k] : 0; CB: bytes per load
template <int TS, int W, int CB>
__global__ void __launch_bounds__(PK_THREADS)
gather_packet_kernel(const uint8_t *__restrict__ source, const SoaPtrs<W> out,
                     const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask, uint64_t n) {
    using R = typename Raw<TS>::type;
    using C = typename Chunk<CB>::type;
    constexpr int PB = TS * W;
    static_assert(PB % CB == 0, "whole chunks");
    const uint64_t stride = (uint64_t) gridDim.x * PK_THREADS;
    for (uint64_t i = (uint64_t) blockIdx.x * PK_THREADS + threadIdx.x; i < n; i += stride) {
        union { R r[W]; C c[PB / CB]; } u;
        #pragma unroll
        for (int k = 0; k < W; ++k)
            u.r[k] = 0;
        if (!mask || __ldcs(mask + i) != 0) {
            const C *p = (const C *) (source + (uint64_t) __ldcs(index + i) * PB);
            #pragma unroll
            for (int q = 0; q < PB / CB; ++q)
                u.c[q] = __ldg(p + q);
        }
        #pragma unroll
        for (int k = 0; k < W; ++k)
            __stcs((R *) out.v[k] + i, u.r[k]);
    }
}

struct PacketIo {
    cudaStream_t stream;
    void *base;              // AoS side
    void *const *soa;        // W device pointers (host array)
    const uint32_t *index;
    const uint8_t *mask;
    uint64_t n;
};

static uint32_t packet_grid(uint64_t n) {
    return (uint32_t) std::max<uint64_t>(
        1, std::min<uint64_t>(ceil_div(n, (uint64_t) PK_THREADS), (uint64_t) sm_count() * 16));
}

template <bool GATHER, int TS, int W, int CB> static int packet_io_launch(const PacketIo &c) {
    SoaPtrs<W> ptrs;
    for (int k = 0; k < W; ++k)
        ptrs.v[k] = c.soa[k];
    if constexpr (GATHER)
        gather_packet_kernel<TS, W, CB><<<packet_grid(c.n), PK_THREADS, 0, c.stream>>>(
            (const uint8_t *) c.base, ptrs, c.index, c.mask, c.n);
    else
        scatter_packet_store_kernel<TS, W, CB><<<packet_grid(c.n), PK_THREADS, 0, c.stream>>>(
            (uint8_t *) c.base, ptrs, c.index, c.mask, c.n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

/// Widest chunk that divides the packet size and the base address
template <bool GATHER, int TS, int W> static int packet_io_chunks(const PacketIo &c) {
    constexpr int PB = TS * W;
    const uintptr_t a = (uintptr_t) c.base;
    if constexpr (PB % 16 == 0)
        if (a % 16 == 0)
            return packet_io_launch<GATHER, TS, W, 16>(c);
    if constexpr (PB % 8 == 0)
        if (a % 8 == 0)
            return packet_io_launch<GATHER, TS, W, 8>(c);
    if constexpr (PB % 4 == 0)
        if (a % 4 == 0)
            return packet_io_launch<GATHER, TS, W, 4>(c);
    if constexpr (PB % 2 == 0)
        if (a % 2 == 0)
            return packet_io_launch<GATHER, TS, W, 2>(c);
    return packet_io_launch<GATHER, TS, W, 1>(c);
}

template <bool GATHER, int TS> static int packet_io_width(const PacketIo &c, uint32_t width) {
    switch (width) {
        case 1: return packet_io_chunks<GATHER, TS, 1>(c);
        case 2: return packet_io_chunks<GATHER, TS, 2>(c);
        case 4: return packet_io_chunks<GATHER, TS, 4>(c);
        case 8: return packet_io_chunks<GATHER, TS, 8>(c);
    }
    return fail(B200_ERR_UNSUPPORTED, "packet size must be 1, 2, 4 or 8 (got %u)", width);
}

template <bool GATHER> static int packet_io(const char *what, void *stream_, int vt, void *base, void *const *soa,
                                            uint32_t width, const uint32_t *index, const uint8_t *mask,
                                            uint64_t n) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (width == 0 || (width & (width - 1)) != 0)
        return fail(B200_ERR_INVALID, "%s(): vector size must be a power of two!", what);
    const uint32_t ts = type_size(vt);
    if (ts == 0)
        return fail(B200_ERR_UNSUPPORTED, "%s(): unsupported variable type (%s)", what, type_name(vt));
    if ((uintptr_t) base % ts != 0)
        return fail(B200_ERR_INVALID, "%s(): misaligned array", what);
    if (n == 0)
        return B200_OK;
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, GATHER ? B200_KERNEL_GATHER : B200_KERNEL_SCATTER, n);
    PacketIo c{ stream, base, soa, index, mask, n };
    switch (ts) {
        case 1: return packet_io_width<GATHER, 1>(c, width);
        case 2: return packet_io_width<GATHER, 2>(c, width);
        case 4: return packet_io_width<GATHER, 4>(c, width);
        default: return packet_io_width<GATHER, 8>(c, width);
    }
}

// ------------------------------------------------------- f16 packet reductions
//
// red.global.v{2,4,8}.f16.add.noftz: one L2 operation per packet of up to 16 bytes
// (the reference emits exactly these for compute capability >= 9.0).  Min / max use
// red.global.v2.f16x2-free scalar forms in scatter.cu; packets of them are issued
// component pair by component pair there.
template <int W> B200_DEVICE void red_add_f16_vec(__half *p, const unsigned short (&h)[W]) {
    if constexpr (W == 8) {
        asm volatile("red.global.v8.f16.add.noftz [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "l"(p), "h"(h[0]), "h"(h[1]), "h"(h[2]), "h"(h[3]), "h"(h[4]), "h"(h[5]), "h"(h[6]), "h"(h[7])
                     : "memory");
    } else if constexpr (W == 4) {
        asm volatile("red.global.v4.f16.add.noftz [%0], {%1, %2, %3, %4};"
                     :: "l"(p), "h"(h[0]), "h"(h[1]), "h"(h[2]), "h"(h[3]) : "memory");
    } else {
        asm volatile("red.global.v2.f16.add.noftz [%0], {%1, %2};" :: "l"(p), "h"(h[0]), "h"(h[1]) : "memory");
    }
}

/// target[index[i] * W + k] += values[k][i] (float16).  MERGE: runs of neighbouring
/// lanes with equal indices are summed first (in float32, one rounding per run).
template <int W, bool MERGE>
__global__ void __launch_bounds__(PK_THREADS)
scatter_packet_f16_add_kernel(__half *__restrict__ target, const SoaPtrs<W> values,
                              const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask, uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t) gridDim.x * PK_THREADS;
    const uint64_t first = (uint64_t) blockIdx.x * PK_THREADS + threadIdx.x;
    for (uint64_t base = first - lane; base < n; base += stride) {
        const uint64_t i = base + lane;
        bool on = i < n;
        uint32_t idx = 0;
        unsigned short h[W];
        #pragma unroll
        for (int k = 0; k < W; ++k)
            h[k] = 0;
        if (on) {
            idx = __ldcs(index + i);
            #pragma unroll
            for (int k = 0; k < W; ++k)
                h[k] = __ldcs((const unsigned short *) values.v[k] + i);
            if (mask)
                on = __ldcs(mask + i) != 0;
        }
        bool issue = on;
        if constexpr (MERGE) {
            const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, idx, 1);
            const uint32_t active = __ballot_sync(FULL_MASK, on);
            const bool head = lane == 0 || idx != prev_idx || !((active >> (lane - 1)) & 1u) || !on;
            const uint32_t heads = __ballot_sync(FULL_MASK, head);
            if (heads != FULL_MASK) {
                const uint32_t above = heads & ~((2u << lane) - 1u);
                const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
                float f[W];
                #pragma unroll
                for (int k = 0; k < W; ++k)
                    f[k] = __half2float(__ushort_as_half(h[k]));
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    #pragma unroll
                    for (int k = 0; k < W; ++k) {
                        const float other = __shfl_sync(FULL_MASK, f[k], min(lane + d, 31u));
                        if (lane + d <= last)
                            f[k] += other;
                    }
                }
                #pragma unroll
                for (int k = 0; k < W; ++k)
                    h[k] = __half_as_ushort(__float2half_rn(f[k]));
            }
            issue = head && on;
        }
        if (issue) {
            __half *p = target + (uint64_t) idx * W;
            if constexpr (W >= 2)
                red_add_f16_vec<W>(p, h);
            else
                atomicAdd(p, __ushort_as_half(h[0]));
        }
    }
}

template <int W> static int packet_f16_add_launch(cudaStream_t stream, __half *target, const void *const *values,
                                                  const uint32_t *index, const uint8_t *mask, uint64_t n,
                                                  bool merge) {
    SoaPtrs<W> ptrs;
    for (int k = 0; k < W; ++k)
        ptrs.v[k] = const_cast<void *>(values[k]);
    if (merge)
        scatter_packet_f16_add_kernel<W, true><<<packet_grid(n), PK_THREADS, 0, stream>>>(target, ptrs, index, mask, n);
    else
        scatter_packet_f16_add_kernel<W, false><<<packet_grid(n), PK_THREADS, 0, stream>>>(target, ptrs, index, mask, n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int packet_f16_add(cudaStream_t stream, void *target, const void *const *values, uint32_t width,
                   const uint32_t *index, const uint8_t *mask, uint64_t n, int mode) {
    const bool merge = mode != B200_MODE_DIRECT;
    __half *t = (__half *) target;
    switch (width) {
        case 1: return packet_f16_add_launch<1>(stream, t, values, index, mask, n, merge);
        case 2: return packet_f16_add_launch<2>(stream, t, values, index, mask, n, merge);
        case 4: return packet_f16_add_launch<4>(stream, t, values, index, mask, n, merge);
        case 8: return packet_f16_add_launch<8>(stream, t, values, index, mask, n, merge);
    }
    return fail(B200_ERR_UNSUPPORTED, "jit_var_scatter_packet(): packet size must be 1, 2, 4 or 8 (got %u)", width);
}

// ------------------------------------- plain scatter, signed / 64-bit indices
//
// Identity: target[index[i]] = value[i] (plain stores; with duplicate indices one of
// the values wins, as in the reference).  Wide indices: the same merge of runs of
// neighbouring equal indices as scatter.cu, on 64-bit element offsets.
template <int TS, typename I>
__global__ void __launch_bounds__(PK_THREADS)
scatter_store_kernel(uint8_t *__restrict__ target, const uint8_t *__restrict__ value,
                     const I *__restrict__ index, const uint8_t *__restrict__ mask, uint64_t n) {
    using R = typename Raw<TS>::type;
    const uint64_t stride = (uint64_t) gridDim.x * PK_THREADS;
    for (uint64_t i = (uint64_t) blockIdx.x * PK_THREADS + threadIdx.x; i < n; i += stride) {
        if (mask && __ldcs(mask + i) == 0)
            continue;
        ((R *) target)[(uint64_t) __ldcs(index + i)] = __ldcs((const R *) value + i);
    }
}

template <typename I> static int scatter_store(cudaStream_t stream, uint32_t ts, void *target, const void *value,
                                               const I *index, const uint8_t *mask, uint64_t n) {
    const uint32_t grid = packet_grid(n);
    uint8_t *t = (uint8_t *) target;
    const uint8_t *v = (const uint8_t *) value;
    switch (ts) {
        case 1: scatter_store_kernel<1, I><<<grid, PK_THREADS, 0, stream>>>(t, v, index, mask, n); break;
        case 2: scatter_store_kernel<2, I><<<grid, PK_THREADS, 0, stream>>>(t, v, index, mask, n); break;
        case 4: scatter_store_kernel<4, I><<<grid, PK_THREADS, 0, stream>>>(t, v, index, mask, n); break;
        default: scatter_store_kernel<8, I><<<grid, PK_THREADS, 0, stream>>>(t, v, index, mask, n); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

/// scatter.cu: atomic scatter-reduce with 64-bit element indices
int scatter_reduce_wide(cudaStream_t stream, int vt, int op, int mode, void *target, const void *value,
                        const uint64_t *index, const uint8_t *mask, uint64_t n);

} // namespace b200

using namespace b200;

extern "C" {

int b200_scatter_packet(void *stream, int vt, void *target, const void *const *values, uint32_t width,
                        const uint32_t *index, const uint8_t *mask, uint64_t n) {
    return packet_io<false>("jit_var_scatter_packet", stream, vt, target, (void *const *) values, width, index,
                            mask, n);
}

int b200_gather_packet(void *stream, int vt, const void *source, void *const *out, uint32_t width,
                       const uint32_t *index, const uint8_t *mask, uint64_t n) {
    return packet_io<true>("jit_var_gather_packet", stream, vt, const_cast<void *>(source), out, width, index,
                           mask, n);
}

int b200_scatter_reduce_idx(void *stream_, int vt, int op, int mode, void *target, const void *value,
                            const void *index, int index_vt, const uint8_t *mask, uint64_t n) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (index_vt != B200_VT_UINT32 && index_vt != B200_VT_INT32 && index_vt != B200_VT_UINT64 &&
        index_vt != B200_VT_INT64)
        return fail(B200_ERR_INVALID, "jit_var_scatter(): the index must be a 32 or 64 bit integer array (%s)!",
                    type_name(index_vt));
    const bool wide = index_vt == B200_VT_UINT64 || index_vt == B200_VT_INT64;
    if (op == B200_OP_IDENTITY) {
        const uint32_t ts = type_size(vt);
        if (ts == 0)
            return fail(B200_ERR_UNSUPPORTED, "jit_var_scatter(): unsupported variable type (%s)", type_name(vt));
        if (n == 0)
            return B200_OK;
        cudaStream_t stream = resolve_stream(stream_);
        HistoryScope hs(stream, B200_KERNEL_SCATTER, n);
        // (valid signed indices are non-negative: same bits as the unsigned type)
        if (wide)
            return scatter_store<uint64_t>(stream, ts, target, value, (const uint64_t *) index, mask, n);
        return scatter_store<uint32_t>(stream, ts, target, value, (const uint32_t *) index, mask, n);
    }
    if (!wide)
        return b200_scatter_reduce(stream_, vt, op, mode, target, value, (const uint32_t *) index, mask, n);
    if (n == 0 || !b200_can_scatter_reduce(vt, op)) // (the uint32 entry point validates and raises)
        return b200_scatter_reduce(stream_, vt, op, mode, target, value, nullptr, mask, 0);
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, B200_KERNEL_SCATTER, n);
    return scatter_reduce_wide(stream, vt, op, mode, target, value, (const uint64_t *) index, mask, n);
}

} // extern "C"
