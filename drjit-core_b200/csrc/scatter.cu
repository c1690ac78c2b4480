// scatter.cu -- atomic scatter-reduce: target[index[i]] op= value[i].
//
// Stand-alone form of what the reference splices into its JIT-generated fused
// kernels as PTX text: jitc_cuda_render_scatter_reduce
// (src/cuda_scatter.cpp:246-354), the warp pre-reduction
// jitc_cuda_render_warp_reduce (:125-244), float min/max emulation through
// signed/unsigned integer atomics (:74-106) and the f16x2 packing trick for
// half precision (:283-339).
//
// Modes (jit.h:1017-1066):
//   Direct       one red.global per element;
//   Local        lanes that hit the same address are merged before the atomic.
//                The reference finds them with match.any, which on B200 costs
//                256 issue cycles per warp instruction whenever the lanes disagree
//                (tools/ubench_warp_ops.cu) -- more than the 32 atomics it can
//                save.  Here a warp picks one of two mergers per batch group from
//                a periodic probe:
//                  runs   neighbouring lanes with equal addresses (what coherent
//                         scatters produce) are merged by a segmented shuffle
//                         reduction -- one shuffle-up, one compare, one ballot to
//                         find the runs; a butterfly when the whole warp agrees;
//                  match  match.any + rank-halving tree, kept for the case where
//                         many duplicates are NOT adjacent (few distinct targets);
//   Auto         Local, except that a warp whose probe shows (almost) no
//                duplicates issues its atomics directly;
//   NoConflicts  plain load / op / store.
#include "common.cuh"

namespace b200 {

static constexpr int SCATTER_THREADS = 256;

// ---- atomic primitives (result unused -> RED instructions) ------------------

template <typename T, int Op> struct Atomic;

#define B200_ATOMIC_INT(T, CT)                                                                  \
    template <> struct Atomic<T, B200_OP_ADD> { static B200_DEVICE void apply(T *p, T v) { atomicAdd((CT *) p, (CT) v); } }; \
    template <> struct Atomic<T, B200_OP_MIN> { static B200_DEVICE void apply(T *p, T v) { atomicMin(p, v); } }; \
    template <> struct Atomic<T, B200_OP_MAX> { static B200_DEVICE void apply(T *p, T v) { atomicMax(p, v); } }; \
    template <> struct Atomic<T, B200_OP_AND> { static B200_DEVICE void apply(T *p, T v) { atomicAnd((CT *) p, (CT) v); } }; \
    template <> struct Atomic<T, B200_OP_OR>  { static B200_DEVICE void apply(T *p, T v) { atomicOr((CT *) p, (CT) v); } };

B200_ATOMIC_INT(uint32_t, unsigned int)
B200_ATOMIC_INT(int32_t, unsigned int)
B200_ATOMIC_INT(unsigned long long, unsigned long long)
B200_ATOMIC_INT(long long, unsigned long long)

template <> struct Atomic<float, B200_OP_ADD> { static B200_DEVICE void apply(float *p, float v) { atomicAdd(p, v); } };
template <> struct Atomic<double, B200_OP_ADD> { static B200_DEVICE void apply(double *p, double v) { atomicAdd(p, v); } };

// Float min/max: IEEE bit patterns order like signed integers when the sign
// bit is clear and like unsigned integers in reverse when it is set.
template <> struct Atomic<float, B200_OP_MIN> {
    static B200_DEVICE void apply(float *p, float v) {
        int bits = __float_as_int(v);
        if (bits >= 0) atomicMin((int *) p, bits);
        else atomicMax((unsigned int *) p, (unsigned int) bits);
    }
};
template <> struct Atomic<float, B200_OP_MAX> {
    static B200_DEVICE void apply(float *p, float v) {
        int bits = __float_as_int(v);
        if (bits >= 0) atomicMax((int *) p, bits);
        else atomicMin((unsigned int *) p, (unsigned int) bits);
    }
};
template <> struct Atomic<double, B200_OP_MIN> {
    static B200_DEVICE void apply(double *p, double v) {
        long long bits = __double_as_longlong(v);
        if (bits >= 0) atomicMin((long long *) p, bits);
        else atomicMax((unsigned long long *) p, (unsigned long long) bits);
    }
};
template <> struct Atomic<double, B200_OP_MAX> {
    static B200_DEVICE void apply(double *p, double v) {
        long long bits = __double_as_longlong(v);
        if (bits >= 0) atomicMax((long long *) p, bits);
        else atomicMin((unsigned long long *) p, (unsigned long long) bits);
    }
};

// Half precision: there is no scalar f16 atomic; operate on the enclosing
// aligned f16x2 word with the identity in the other half.
template <int Op> B200_DEVICE void atomic_f16x2(__half *p, __half v, unsigned short identity) {
    uintptr_t addr = (uintptr_t) p;
    bool upper = (addr & 2) != 0;
    unsigned short vb = __half_as_ushort(v);
    unsigned short lo = upper ? identity : vb, hi = upper ? vb : identity;
    void *word = (void *) (addr & ~(uintptr_t) 2);
    if constexpr (Op == B200_OP_ADD) {
        uint32_t packed = ((uint32_t) hi << 16) | lo;
        asm volatile("red.global.add.noftz.f16x2 [%0], %1;" :: "l"(word), "r"(packed) : "memory");
    } else if constexpr (Op == B200_OP_MIN) {
        // sm_90+ vector form (the reference emits the same, cuda_scatter.cpp:307-332)
        asm volatile("red.global.v2.f16.min.noftz [%0], {%1, %2};" :: "l"(word), "h"(lo), "h"(hi) : "memory");
    } else {
        asm volatile("red.global.v2.f16.max.noftz [%0], {%1, %2};" :: "l"(word), "h"(lo), "h"(hi) : "memory");
    }
}
template <> struct Atomic<__half, B200_OP_ADD> { static B200_DEVICE void apply(__half *p, __half v) { atomic_f16x2<B200_OP_ADD>(p, v, 0x0000); } };
template <> struct Atomic<__half, B200_OP_MIN> { static B200_DEVICE void apply(__half *p, __half v) { atomic_f16x2<B200_OP_MIN>(p, v, 0x7c00); } };
template <> struct Atomic<__half, B200_OP_MAX> { static B200_DEVICE void apply(__half *p, __half v) { atomic_f16x2<B200_OP_MAX>(p, v, 0xfc00); } };

// Operator in the element type itself (the pre-reduction of the reference runs
// in the element type as well, half included: src/cuda_scatter.cpp:133-135)
template <typename T, int Op> struct ElemOp {
    static B200_DEVICE T apply(T a, T b) { return Red<T, Op>::apply(a, b); }
};
template <int Op> struct ElemOp<__half, Op> {
    static B200_DEVICE __half apply(__half a, __half b) {
        if constexpr (Op == B200_OP_ADD) return __hadd(a, b);
        else if constexpr (Op == B200_OP_MIN) return __hmin(a, b);
        else return __hmax(a, b);
    }
};

template <typename T> B200_DEVICE T shfl_elem(uint32_t mask, T v, int src) {
    if constexpr (sizeof(T) == 8) {
        uint64_t u; memcpy(&u, &v, 8);
        uint32_t lo = __shfl_sync(mask, (uint32_t) u, src), hi = __shfl_sync(mask, (uint32_t) (u >> 32), src);
        u = ((uint64_t) hi << 32) | lo;
        T r; memcpy(&r, &u, 8); return r;
    } else if constexpr (sizeof(T) == 4) {
        uint32_t u; memcpy(&u, &v, 4);
        u = __shfl_sync(mask, u, src);
        T r; memcpy(&r, &u, 4); return r;
    } else {
        unsigned short s; memcpy(&s, &v, 2);
        uint32_t u = __shfl_sync(mask, (uint32_t) s, src);
        s = (unsigned short) u;
        T r; memcpy(&r, &s, 2); return r;
    }
}

/// Warp pre-reduction of the lanes in 'active' that hit the same address
/// (src/cuda_scatter.cpp:125-244); the lowest lane of every group issues the
/// single atomic.  Returns the number of lanes that were absorbed by another one.
template <typename T, int Op>
B200_DEVICE uint32_t scatter_local(T *target, uint32_t idx, T v, bool on, uint32_t lane) {
    const uint32_t active = __ballot_sync(FULL_MASK, on);
    uint32_t absorbed = 0;
    if (on) {
        const uint32_t peers = __match_any_sync(active, idx);
        const uint32_t lower = peers & ((1u << lane) - 1);
        if (peers == FULL_MASK) {
            // all 32 lanes hit one address: butterfly
            #pragma unroll
            for (int d = 16; d > 0; d >>= 1)
                v = ElemOp<T, Op>::apply(v, shfl_elem<T>(FULL_MASK, v, lane ^ d));
        } else if (__any_sync(active, peers != (1u << lane))) {
            // rank-halving tree inside every group of equal addresses: each
            // round a lane absorbs the next surviving peer above it, then the
            // odd-ranked lanes drop out
            uint32_t rank = __popc(lower);
            uint32_t above = peers & ~((2u << lane) - 1);
            while (__any_sync(active, above != 0)) {
                int src = above ? __ffs(above) - 1 : (int) lane;
                T other = shfl_elem<T>(active, v, src);
                if (above)
                    v = ElemOp<T, Op>::apply(v, other);
                uint32_t even = __ballot_sync(active, (rank & 1) == 0);
                above &= even;
                rank >>= 1;
            }
        }
        if (lower == 0) // lowest lane of its group
            Atomic<T, Op>::apply(target + idx, v);
        absorbed = __popc(__ballot_sync(active, lower != 0));
    }
    return absorbed;
}

/// Warp pre-reduction of RUNS of neighbouring lanes with equal addresses; the
/// first lane of every run issues the atomic.  Duplicates that are not adjacent
/// stay separate atomics (still correct).
template <typename T, int Op>
B200_DEVICE void scatter_runs(T *target, uint32_t idx, T v, bool on, uint32_t lane) {
    const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, idx, 1);
    const uint32_t active = __ballot_sync(FULL_MASK, on);
    // a lane starts a run unless its lower neighbour is active with the same index
    const bool head = lane == 0 || idx != prev_idx || !((active >> (lane - 1)) & 1u) || !on;
    const uint32_t heads = __ballot_sync(FULL_MASK, head);
    if (heads == 1u) {
        // the whole warp hits one address
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            v = ElemOp<T, Op>::apply(v, shfl_elem<T>(FULL_MASK, v, lane ^ d));
        if (lane == 0)
            Atomic<T, Op>::apply(target + idx, v);
        return;
    }
    if (heads != FULL_MASK) {
        // last lane of this lane's run: one below the next head (inactive lanes are
        // heads of their own, so a run never extends over them)
        const uint32_t above = heads & ~((2u << lane) - 1u);
        const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T other = shfl_elem<T>(FULL_MASK, v, min(lane + d, 31u));
            if (lane + d <= last)
                v = ElemOp<T, Op>::apply(v, other);
        }
    }
    if (head && on)
        Atomic<T, Op>::apply(target + idx, v);
}

/// MODE: Direct / Local / NoConflicts as in jit.h:1017-1066.  Local picks its
/// merger per warp from a probe (every eighth batch, first step): match.any counts
/// all duplicates among the 32 addresses, the run detection the adjacent ones.
///   AUTO and (almost) no duplicates          -> direct atomics
///   most duplicates adjacent (or none)       -> scatter_runs
///   otherwise                                -> scatter_local (match.any)
enum { PATH_DIRECT = 0, PATH_RUNS = 1, PATH_MATCH = 2 };

template <typename T, int Op, int MODE, bool AUTO, int U>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_reduce_kernel(T *__restrict__ target, const T *__restrict__ value,
                      const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask,
                      uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    int path = PATH_RUNS; // warp-uniform
    uint32_t batch = 0;

    // a CTA walks tiles of U * 256 CONSECUTIVE elements (the U loads of a thread stay
    // within one 4 - 8 KiB neighbourhood); all lanes of a warp run the same number of
    // iterations (warp collectives)
    constexpr uint64_t TILE = (uint64_t) SCATTER_THREADS * U;
    for (uint64_t tile0 = (uint64_t) blockIdx.x * TILE; tile0 < n; tile0 += (uint64_t) gridDim.x * TILE, ++batch) {
        const uint64_t base = tile0 + (threadIdx.x & ~31u);
        constexpr uint64_t stride = SCATTER_THREADS;
        T val[U];
        uint32_t idx[U];
        bool on[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t i = base + lane + (uint64_t) u * stride;
            on[u] = i < n;
            idx[u] = 0;
            if (on[u]) {
                idx[u] = __ldcs(index + i);
                val[u] = __ldcs(value + i);
                if (mask)
                    on[u] = __ldcs(mask + i) != 0;
            }
        }
        const bool probe = MODE == B200_MODE_LOCAL && (batch & 7) == 0;
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base + (uint64_t) u * stride >= n)
                break; // warp-uniform
            if constexpr (MODE == B200_MODE_NO_CONFLICTS) {
                if (on[u]) {
                    T *p = target + idx[u];
                    *p = ElemOp<T, Op>::apply(*p, val[u]);
                }
            } else if constexpr (MODE == B200_MODE_DIRECT) {
                if (on[u])
                    Atomic<T, Op>::apply(target + idx[u], val[u]);
            } else {
                if (probe && u == 0) {
                    // adjacent duplicates (what scatter_runs would merge)
                    const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, idx[u], 1);
                    const uint32_t active = __ballot_sync(FULL_MASK, on[u]);
                    const bool cont = lane != 0 && on[u] && idx[u] == prev_idx && ((active >> (lane - 1)) & 1u);
                    const uint32_t adjacent = __popc(__ballot_sync(FULL_MASK, cont));
                    // all duplicates (and the merged atomics of this step)
                    const uint32_t all = scatter_local<T, Op>(target, idx[u], val[u], on[u], lane);
                    const uint32_t dup = __shfl_sync(FULL_MASK, all, __ffs(active | 0x80000000u) - 1);
                    if (AUTO && dup < 4)
                        path = PATH_DIRECT;
                    else if (4 * (dup - adjacent) <= dup)
                        path = PATH_RUNS;
                    else
                        path = PATH_MATCH;
                } else if (path == PATH_RUNS) {
                    scatter_runs<T, Op>(target, idx[u], val[u], on[u], lane);
                } else if (path == PATH_MATCH) {
                    scatter_local<T, Op>(target, idx[u], val[u], on[u], lane);
                } else if (on[u]) {
                    Atomic<T, Op>::apply(target + idx[u], val[u]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- scatter_inc
//
// jit_var_scatter_inc (jit.h:1125-1143; emitter src/cuda_scatter.cpp:356-393):
// atomically increment target[index[i]] and return the OLD value -- the building
// block of queue-style stream compaction.  Warp aggregated like the reference (one
// atomic per group of lanes with the same counter, every lane gets base + rank),
// but the groups are found without match.any: a vote when the whole warp hits one
// counter (the compaction case), runs of neighbouring equal indices otherwise.
// Masked lanes receive 0 (src/cuda_scatter.cpp:361-364).
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_inc_kernel(uint32_t *__restrict__ target, const uint32_t *__restrict__ index,
                   const uint8_t *__restrict__ mask, uint32_t *__restrict__ out, uint64_t n) {
    constexpr int WARPS = SCATTER_THREADS / 32;
    __shared__ uint32_t s_idx[WARPS], s_cnt[WARPS], s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t stride = (uint64_t) gridDim.x * SCATTER_THREADS;
    // CTA-uniform trip count (the loop contains block-wide barriers)
    for (uint64_t cbase = (uint64_t) blockIdx.x * SCATTER_THREADS; cbase < n; cbase += stride) {
        const uint64_t i = cbase + threadIdx.x;
        bool on = i < n;
        uint32_t idx = 0;
        if (on) {
            idx = __ldcs(index + i);
            if (mask)
                on = __ldcs(mask + i) != 0;
        }
        const uint32_t active = __ballot_sync(FULL_MASK, on);
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t lead = __ffs(active | 0x80000000u) - 1;
        const uint32_t lead_idx = __shfl_sync(FULL_MASK, idx, lead);
        const bool warp_uniform = __all_sync(FULL_MASK, !on || idx == lead_idx);

        // Whole CTA on one counter (queue compaction): ONE atomic per 256 entries --
        // atomics on a single address are served one at a time by the L2 (measured:
        // 2^21 warp-level atomics on one counter take 1.4 ms)
        if (lane == 0) {
            s_cnt[warp] = __popc(active);
            s_idx[warp] = warp_uniform ? lead_idx : 0xffffffffu; // (index 2^32 - 1 takes the warp path)
        }
        __syncthreads();
        uint32_t total = 0, before = 0, ref = 0xffffffffu;
        bool cta_uniform = true;
        #pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_cnt[w], x = s_idx[w];
            if (c) {
                if (ref == 0xffffffffu)
                    ref = x;
                cta_uniform &= x == ref && x != 0xffffffffu;
            }
            total += c;
            before += (uint32_t) w < warp ? c : 0u;
        }
        uint32_t old = 0;
        if (cta_uniform && total) {
            if (threadIdx.x == 0)
                s_base = atomicAdd(target + ref, total);
            __syncthreads();
            old = s_base + before + __popc(active & lt);
        } else if (active) {
            if (warp_uniform) {
                uint32_t prev = 0;
                if (lane == lead)
                    prev = atomicAdd(target + idx, (uint32_t) __popc(active));
                prev = __shfl_sync(FULL_MASK, prev, lead);
                old = prev + __popc(active & lt);
            } else {
                const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, idx, 1);
                const bool head = lane == 0 || idx != prev_idx || !((active >> (lane - 1)) & 1u) || !on;
                const uint32_t heads = __ballot_sync(FULL_MASK, head);
                // this lane's run: [my_head, last]
                const uint32_t my_head = 31u - __clz(heads & (lt | (1u << lane)));
                const uint32_t above = heads & ~((2u << lane) - 1u);
                const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
                uint32_t prev = 0;
                if (head && on)
                    prev = atomicAdd(target + idx, last - lane + 1u);
                prev = __shfl_sync(FULL_MASK, prev, my_head);
                old = prev + (lane - my_head);
            }
        }
        if (i < n)
            out[i] = on ? old : 0u;
        __syncthreads(); // s_idx / s_cnt / s_base are rewritten in the next step
    }
}

/// Four consecutive entries per thread (128-bit loads / stores; 16-byte aligned
/// index and out arrays).  A thread whose active entries all name one counter
/// takes part as ONE unit with a count of up to four: the CTA-wide path then
/// issues one atomic per 1024 entries (one shared counter: 0.44 -> 0.16 ms for
/// 2^26 entries), the warp path one per run of neighbouring threads with the
/// same counter.  Warps in which some thread names several counters issue one
/// atomic per entry.
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_inc_vec4_kernel(uint32_t *__restrict__ target, const uint32_t *__restrict__ index,
                        const uint8_t *__restrict__ mask, uint32_t *__restrict__ out, uint64_t nvec) {
    constexpr int WARPS = SCATTER_THREADS / 32;
    __shared__ uint32_t s_idx[WARPS], s_cnt[WARPS], s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t stride = (uint64_t) gridDim.x * SCATTER_THREADS;
    for (uint64_t cbase = (uint64_t) blockIdx.x * SCATTER_THREADS; cbase < nvec; cbase += stride) {
        const uint64_t q = cbase + threadIdx.x;
        const bool in_range = q < nvec;
        uint32_t idx[4] = { 0, 0, 0, 0 };
        bool on[4] = { false, false, false, false };
        if (in_range) {
            const uint4 i4 = ld_stream(index + 4 * q);
            idx[0] = i4.x; idx[1] = i4.y; idx[2] = i4.z; idx[3] = i4.w;
            uint32_t m4 = 0x01010101u;
            if (mask)
                m4 = __ldcs((const uint32_t *) (mask + 4 * q));
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                on[k] = ((m4 >> (8 * k)) & 0xffu) != 0;
        }
        // the thread as a unit: number of active entries, their common counter (if any)
        uint32_t t_cnt = 0, t_idx = 0;
        bool t_uniform = true;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (on[k]) {
                if (t_cnt == 0)
                    t_idx = idx[k];
                else
                    t_uniform &= idx[k] == t_idx;
                t_cnt++;
            }
        }
        const uint32_t active = __ballot_sync(FULL_MASK, t_cnt != 0);
        const bool units = __all_sync(FULL_MASK, t_uniform); // warp-uniform
        const uint32_t lead = __ffs(active | 0x80000000u) - 1;
        const uint32_t lead_idx = __shfl_sync(FULL_MASK, t_idx, lead);
        const bool warp_uniform = units && __all_sync(FULL_MASK, t_cnt == 0 || t_idx == lead_idx);
        // inclusive prefix of the unit counts over the lanes
        uint32_t incl = t_cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= (uint32_t) d)
                incl += up;
        }
        const uint32_t warp_total = __shfl_sync(FULL_MASK, incl, 31);

        if (lane == 0) {
            s_cnt[warp] = warp_total;
            s_idx[warp] = warp_uniform ? lead_idx : 0xffffffffu;
        }
        __syncthreads();
        uint32_t total = 0, before = 0, ref = 0xffffffffu;
        bool cta_uniform = true;
        #pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_cnt[w], x = s_idx[w];
            if (c) {
                if (ref == 0xffffffffu)
                    ref = x;
                cta_uniform &= x == ref && x != 0xffffffffu;
            }
            total += c;
            before += (uint32_t) w < warp ? c : 0u;
        }

        uint32_t old[4] = { 0, 0, 0, 0 };
        if (cta_uniform && total) {
            if (threadIdx.x == 0)
                s_base = atomicAdd(target + ref, total);
            __syncthreads();
            uint32_t o = s_base + before + incl - t_cnt;
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                if (on[k])
                    old[k] = o++;
        } else if (units) {
            if (active) {
                // runs of neighbouring threads with the same counter
                const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, t_idx, 1);
                const bool head = lane == 0 || t_cnt == 0 || !((active >> (lane - 1)) & 1u) || t_idx != prev_idx;
                const uint32_t heads = __ballot_sync(FULL_MASK, head);
                const uint32_t my_head = 31u - __clz(heads & (lt | (1u << lane)));
                const uint32_t above = heads & ~((2u << lane) - 1u);
                const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
                // entries in this lane's run: inclusive prefix at its end minus exclusive at its start
                const uint32_t head_excl = __shfl_sync(FULL_MASK, incl - t_cnt, my_head);
                const uint32_t last_incl = __shfl_sync(FULL_MASK, incl, last);
                uint32_t prev = 0;
                if (head && t_cnt)
                    prev = atomicAdd(target + t_idx, last_incl - head_excl);
                prev = __shfl_sync(FULL_MASK, prev, my_head);
                uint32_t o = prev + (incl - t_cnt) - head_excl;
                #pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (on[k])
                        old[k] = o++;
            }
        } else {
            // some thread names several counters (incoherent indices): one atomic per
            // entry, all four of a thread in flight together
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                if (on[k])
                    old[k] = atomicAdd(target + idx[k], 1u);
        }
        if (in_range) {
            uint4 o4;
            o4.x = on[0] ? old[0] : 0u; o4.y = on[1] ? old[1] : 0u;
            o4.z = on[2] ? old[2] : 0u; o4.w = on[3] ? old[3] : 0u;
            st_stream(out + 4 * q, o4);
        }
        __syncthreads(); // s_idx / s_cnt / s_base are rewritten in the next step
    }
}

// -------------------------------------------------------------- packet scatter
//
// jit_var_scatter_packet with a reduction (jit.h:1117; emitter
// jitc_cuda_render_scatter_reduce_packet, src/cuda_packet.cpp:169-327): every
// element carries W (power of two) consecutive components,
//   target[index[i] * W + k] op= values[k][i],   k < W,
// i.e. an array-of-structures target (RGB(A) splats) fed from W separate arrays.
// f32 Add uses the vector reductions of sm_90+ (red.global.add.v4.f32 / .v2.f32,
// as the reference does for cc >= 90, cuda_packet.cpp:243-289): one L2 operation
// per 16 bytes instead of four.  Everything else is W scalar atomics.  In Local /
// Auto mode runs of neighbouring equal indices are merged first (component-wise
// segmented shuffle reduction, as in scatter_runs).
template <typename T, int W> struct PacketPtrs { const T *v[W]; };

template <int W> B200_DEVICE void red_add_f32_vec(float *p, const float (&v)[W]) {
    if constexpr (W % 4 == 0) {
        #pragma unroll
        for (int k = 0; k < W; k += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         :: "l"(p + k), "f"(v[k]), "f"(v[k + 1]), "f"(v[k + 2]), "f"(v[k + 3]) : "memory");
    } else if constexpr (W == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        atomicAdd(p, v[0]);
    }
}

template <typename T, int Op, int W, bool MERGE>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_packet_kernel(T *__restrict__ target, const PacketPtrs<T, W> values,
                      const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask,
                      uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t) gridDim.x * SCATTER_THREADS;
    const uint64_t first = (uint64_t) blockIdx.x * SCATTER_THREADS + threadIdx.x;
    for (uint64_t base = first - lane; base < n; base += stride) {
        const uint64_t i = base + lane;
        bool on = i < n;
        uint32_t idx = 0;
        T v[W];
        if (on) {
            idx = __ldcs(index + i);
            #pragma unroll
            for (int k = 0; k < W; ++k)
                v[k] = __ldcs(values.v[k] + i);
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
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    #pragma unroll
                    for (int k = 0; k < W; ++k) {
                        const T other = shfl_elem<T>(FULL_MASK, v[k], min(lane + d, 31u));
                        if (lane + d <= last)
                            v[k] = ElemOp<T, Op>::apply(v[k], other);
                    }
                }
            }
            issue = head && on;
        }
        if (issue) {
            T *p = target + (uint64_t) idx * W;
            if constexpr (std::is_same<T, float>::value && Op == B200_OP_ADD && W >= 2) {
                red_add_f32_vec<W>(p, v);
            } else {
                #pragma unroll
                for (int k = 0; k < W; ++k)
                    Atomic<T, Op>::apply(p + k, v[k]);
            }
        }
    }
}

struct PacketCall {
    cudaStream_t stream;
    void *target;
    const void *const *values;
    uint32_t width;
    const uint32_t *index;
    const uint8_t *mask;
    uint64_t n;
    int mode;
};

template <typename T, int Op, int W> static int launch_packet_w(const PacketCall &c) {
    PacketPtrs<T, W> ptrs;
    for (int k = 0; k < W; ++k)
        ptrs.v[k] = (const T *) c.values[k];
    uint32_t grid = (uint32_t) std::max<uint64_t>(
        1, std::min<uint64_t>(ceil_div(c.n, (uint64_t) SCATTER_THREADS), (uint64_t) sm_count() * 16));
    if (c.mode == B200_MODE_DIRECT)
        scatter_packet_kernel<T, Op, W, false><<<grid, SCATTER_THREADS, 0, c.stream>>>(
            (T *) c.target, ptrs, c.index, c.mask, c.n);
    else
        scatter_packet_kernel<T, Op, W, true><<<grid, SCATTER_THREADS, 0, c.stream>>>(
            (T *) c.target, ptrs, c.index, c.mask, c.n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <typename T, int Op> static int launch_packet(const PacketCall &c) {
    switch (c.width) {
        case 1: return launch_packet_w<T, Op, 1>(c);
        case 2: return launch_packet_w<T, Op, 2>(c);
        case 4: return launch_packet_w<T, Op, 4>(c);
        case 8: return launch_packet_w<T, Op, 8>(c);
        default:
            return fail(B200_ERR_UNSUPPORTED,
                        "jit_var_scatter_packet(): packet size must be 1, 2, 4 or 8 (got %u)", c.width);
    }
}

typedef int (*PacketFn)(const PacketCall &);

static PacketFn pick_packet(int vt, int op) {
    // the types / operations splatting uses; everything else: W calls of b200_scatter_reduce
    if (vt == B200_VT_FLOAT32) {
        switch (op) {
            case B200_OP_ADD: return launch_packet<float, B200_OP_ADD>;
            case B200_OP_MIN: return launch_packet<float, B200_OP_MIN>;
            case B200_OP_MAX: return launch_packet<float, B200_OP_MAX>;
        }
    } else if (vt == B200_VT_UINT32) {
        switch (op) {
            case B200_OP_ADD: return launch_packet<uint32_t, B200_OP_ADD>;
            case B200_OP_MIN: return launch_packet<uint32_t, B200_OP_MIN>;
            case B200_OP_MAX: return launch_packet<uint32_t, B200_OP_MAX>;
            case B200_OP_AND: return launch_packet<uint32_t, B200_OP_AND>;
            case B200_OP_OR:  return launch_packet<uint32_t, B200_OP_OR>;
        }
    } else if (vt == B200_VT_INT32) {
        switch (op) {
            case B200_OP_ADD: return launch_packet<int32_t, B200_OP_ADD>;
            case B200_OP_MIN: return launch_packet<int32_t, B200_OP_MIN>;
            case B200_OP_MAX: return launch_packet<int32_t, B200_OP_MAX>;
        }
    } else if (vt == B200_VT_FLOAT64 && op == B200_OP_ADD) {
        return launch_packet<double, B200_OP_ADD>;
    }
    return nullptr;
}

/// Local / Auto for 4-byte element types with 16-byte aligned inputs: a lane loads
/// FOUR consecutive elements (one 128-bit load per array), merges equal neighbours
/// among its own four first, and only lanes whose four elements all hit one address
/// take part in the warp-level run merge -- one shuffle round per 128 elements
/// instead of per 32 (the merge is shuffle-pipe bound: ~46 cycles per round and
/// sub-partition).  Incoherent data skip the warp stage altogether.  A periodic
/// probe (match.any over one component) switches to scatter_local when duplicates
/// are frequent but not adjacent (few distinct targets).
template <typename T, int Op>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_reduce_vec4_kernel(T *__restrict__ target, const T *__restrict__ value,
                           const uint32_t *__restrict__ index, const uint8_t *__restrict__ mask,
                           uint64_t nvec) {
    static_assert(sizeof(T) == 4, "four elements per 16-byte load");
    const uint32_t lane = threadIdx.x & 31;
    bool use_match = false; // warp-uniform
    uint32_t batch = 0;
    for (uint64_t tile0 = (uint64_t) blockIdx.x * SCATTER_THREADS; tile0 < nvec;
         tile0 += (uint64_t) gridDim.x * SCATTER_THREADS, ++batch) {
        const uint64_t q = tile0 + threadIdx.x; // vector index (warp-uniform trip count)
        const bool in_range = q < nvec;
        uint32_t idx[4] = { 0, 0, 0, 0 };
        T val[4];
        bool on[4] = { false, false, false, false };
        if (in_range) {
            const uint4 i4 = ld_stream(index + 4 * q);
            const uint4 v4 = ld_stream(value + 4 * q);
            idx[0] = i4.x; idx[1] = i4.y; idx[2] = i4.z; idx[3] = i4.w;
            memcpy(&val[0], &v4.x, 4); memcpy(&val[1], &v4.y, 4);
            memcpy(&val[2], &v4.z, 4); memcpy(&val[3], &v4.w, 4);
            uint32_t m4 = 0x01010101u;
            if (mask)
                m4 = __ldcs((const uint32_t *) (mask + 4 * q));
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                on[k] = ((m4 >> (8 * k)) & 0xffu) != 0;
        }
        const bool single = on[0] && on[1] && on[2] && on[3] && idx[0] == idx[1] && idx[1] == idx[2] &&
                            idx[2] == idx[3];
        const uint32_t singles = __ballot_sync(FULL_MASK, single);

        if ((batch & 7) == 0) {
            // probe: duplicates among the first components that lane coherence does not explain
            const uint32_t active = __ballot_sync(FULL_MASK, on[0]);
            uint32_t leaders = 0;
            if (on[0])
                leaders = (__match_any_sync(active, idx[0]) & ((1u << lane) - 1u)) == 0;
            const uint32_t distinct = __popc(__ballot_sync(FULL_MASK, leaders != 0));
            const uint32_t dup = __popc(active) - distinct;
            use_match = dup >= 8 && __popc(singles) < 16;
        }
        if (use_match) {
            #pragma unroll
            for (int k = 0; k < 4; ++k)
                scatter_local<T, Op>(target, idx[k], val[k], on[k], lane);
            continue;
        }

        if (single) {
            // the lane's four elements: one value
            T v = ElemOp<T, Op>::apply(ElemOp<T, Op>::apply(val[0], val[1]), ElemOp<T, Op>::apply(val[2], val[3]));
            val[0] = v;
        }
        if (singles) { // warp-uniform: runs of neighbouring single lanes with equal index
            const uint32_t prev_idx = __shfl_up_sync(FULL_MASK, idx[0], 1);
            const bool head = lane == 0 || !single || !((singles >> (lane - 1)) & 1u) || idx[0] != prev_idx;
            const uint32_t heads = __ballot_sync(FULL_MASK, head);
            if (heads != FULL_MASK) {
                const uint32_t above = heads & ~((2u << lane) - 1u);
                const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
                T v = val[0];
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const T other = shfl_elem<T>(FULL_MASK, v, min(lane + d, 31u));
                    if (lane + d <= last)
                        v = ElemOp<T, Op>::apply(v, other);
                }
                val[0] = v;
            }
            if (single) {
                if (head)
                    Atomic<T, Op>::apply(target + idx[0], val[0]);
                continue; // (lane-divergent: nothing warp-wide follows)
            }
        }
        // mixed lane: merge equal neighbours among the four, one atomic per segment
        T acc = val[0];
        uint32_t cur = idx[0];
        bool cur_on = on[0];
        #pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (on[k] && cur_on && idx[k] == cur) {
                acc = ElemOp<T, Op>::apply(acc, val[k]);
            } else {
                if (cur_on)
                    Atomic<T, Op>::apply(target + cur, acc);
                acc = val[k];
                cur = idx[k];
                cur_on = on[k];
            }
        }
        if (cur_on)
            Atomic<T, Op>::apply(target + cur, acc);
    }
}

struct ScatterCall {
    cudaStream_t stream;
    void *target;
    const void *value;
    const uint32_t *index;
    const uint8_t *mask;
    uint64_t n;
    int mode;
};

template <typename T, int Op> static int launch_scatter(const ScatterCall &c) {
    constexpr int U = 4;
    uint32_t grid = (uint32_t) std::max<uint64_t>(
        1, std::min<uint64_t>(ceil_div(c.n, (uint64_t) SCATTER_THREADS * U), (uint64_t) sm_count() * 16));
    T *target = (T *) c.target;
    const T *value = (const T *) c.value;
    if constexpr (sizeof(T) == 4) {
        // Local / Auto, 16-byte aligned inputs: four elements per lane; the (< 4) tail
        // elements go through the scalar kernel
        if ((c.mode == B200_MODE_LOCAL || c.mode == B200_MODE_AUTO) && c.n >= 4096 &&
            ((uintptr_t) c.index % 16) == 0 && ((uintptr_t) c.value % 16) == 0 &&
            (!c.mask || ((uintptr_t) c.mask % 4) == 0)) {
            const uint64_t nvec = c.n / 4;
            uint32_t vgrid = (uint32_t) std::max<uint64_t>(
                1, std::min<uint64_t>(ceil_div(nvec, (uint64_t) SCATTER_THREADS), (uint64_t) sm_count() * 16));
            scatter_reduce_vec4_kernel<T, Op><<<vgrid, SCATTER_THREADS, 0, c.stream>>>(target, value, c.index,
                                                                                     c.mask, nvec);
            B200_LAUNCH_CHECK();
            if (c.n % 4 == 0)
                return B200_OK;
            ScatterCall tail = c;
            tail.value = value + 4 * nvec;
            tail.index = c.index + 4 * nvec;
            tail.mask = c.mask ? c.mask + 4 * nvec : nullptr;
            tail.n = c.n % 4;
            tail.mode = B200_MODE_DIRECT;
            return launch_scatter<T, Op>(tail);
        }
    }
    switch (c.mode) {
        case B200_MODE_DIRECT:
            scatter_reduce_kernel<T, Op, B200_MODE_DIRECT, false, U><<<grid, SCATTER_THREADS, 0, c.stream>>>(
                target, value, c.index, c.mask, c.n);
            break;
        case B200_MODE_NO_CONFLICTS:
            scatter_reduce_kernel<T, Op, B200_MODE_NO_CONFLICTS, false, U><<<grid, SCATTER_THREADS, 0, c.stream>>>(
                target, value, c.index, c.mask, c.n);
            break;
        case B200_MODE_LOCAL:
            scatter_reduce_kernel<T, Op, B200_MODE_LOCAL, false, U><<<grid, SCATTER_THREADS, 0, c.stream>>>(
                target, value, c.index, c.mask, c.n);
            break;
        default: // Auto
            scatter_reduce_kernel<T, Op, B200_MODE_LOCAL, true, U><<<grid, SCATTER_THREADS, 0, c.stream>>>(
                target, value, c.index, c.mask, c.n);
            break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

typedef int (*ScatterFn)(const ScatterCall &);

template <typename T> static ScatterFn pick_scatter_int(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_scatter<T, B200_OP_ADD>;
        case B200_OP_MIN: return launch_scatter<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_scatter<T, B200_OP_MAX>;
        case B200_OP_AND: return launch_scatter<T, B200_OP_AND>;
        case B200_OP_OR:  return launch_scatter<T, B200_OP_OR>;
        default: return nullptr;
    }
}

template <typename T> static ScatterFn pick_scatter_float(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_scatter<T, B200_OP_ADD>;
        case B200_OP_MIN: return launch_scatter<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_scatter<T, B200_OP_MAX>;
        default: return nullptr;
    }
}

static ScatterFn pick_scatter(int vt, int op) {
    switch (vt) {
        case B200_VT_INT32:   return pick_scatter_int<int32_t>(op);
        case B200_VT_UINT32:  return pick_scatter_int<uint32_t>(op);
        case B200_VT_INT64:   return pick_scatter_int<long long>(op);
        case B200_VT_UINT64:  return pick_scatter_int<unsigned long long>(op);
        case B200_VT_FLOAT16: return pick_scatter_float<__half>(op);
        case B200_VT_FLOAT32: return pick_scatter_float<float>(op);
        case B200_VT_FLOAT64: return pick_scatter_float<double>(op);
        default: return nullptr;
    }
}


// ----------------------------------------------------------- 64-bit indices
//
// jitc_var_scatter accepts uint64 / int64 indices as well (src/op.cpp:2899-3086).
// Same structure as the scalar kernel with the run merge of scatter_runs, on 64-bit
// element offsets (valid signed indices are non-negative: same bits as unsigned).
template <typename T, int Op, bool MERGE>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_reduce_wide_kernel(T *__restrict__ target, const T *__restrict__ value,
                           const uint64_t *__restrict__ index, const uint8_t *__restrict__ mask,
                           uint64_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t) gridDim.x * SCATTER_THREADS;
    const uint64_t first = (uint64_t) blockIdx.x * SCATTER_THREADS + threadIdx.x;
    for (uint64_t base = first - lane; base < n; base += stride) {
        const uint64_t i = base + lane;
        bool on = i < n;
        uint64_t idx = 0;
        T v = T();
        if (on) {
            idx = __ldcs(index + i);
            v = __ldcs(value + i);
            if (mask)
                on = __ldcs(mask + i) != 0;
        }
        bool issue = on;
        if constexpr (MERGE) {
            const uint64_t prev_idx = __shfl_up_sync(FULL_MASK, idx, 1);
            const uint32_t active = __ballot_sync(FULL_MASK, on);
            const bool head = lane == 0 || idx != prev_idx || !((active >> (lane - 1)) & 1u) || !on;
            const uint32_t heads = __ballot_sync(FULL_MASK, head);
            if (heads != FULL_MASK) {
                const uint32_t above = heads & ~((2u << lane) - 1u);
                const uint32_t last = above ? (uint32_t) __ffs(above) - 2u : 31u;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const T other = shfl_elem<T>(FULL_MASK, v, min(lane + d, 31u));
                    if (lane + d <= last)
                        v = ElemOp<T, Op>::apply(v, other);
                }
            }
            issue = head && on;
        }
        if (issue)
            Atomic<T, Op>::apply(target + idx, v);
    }
}

struct WideCall {
    cudaStream_t stream;
    void *target;
    const void *value;
    const uint64_t *index;
    const uint8_t *mask;
    uint64_t n;
    int mode;
};

template <typename T, int Op> static int launch_scatter_wide(const WideCall &c) {
    uint32_t grid = (uint32_t) std::max<uint64_t>(
        1, std::min<uint64_t>(ceil_div(c.n, (uint64_t) SCATTER_THREADS), (uint64_t) sm_count() * 16));
    if (c.mode == B200_MODE_DIRECT || c.mode == B200_MODE_NO_CONFLICTS)
        scatter_reduce_wide_kernel<T, Op, false><<<grid, SCATTER_THREADS, 0, c.stream>>>(
            (T *) c.target, (const T *) c.value, c.index, c.mask, c.n);
    else
        scatter_reduce_wide_kernel<T, Op, true><<<grid, SCATTER_THREADS, 0, c.stream>>>(
            (T *) c.target, (const T *) c.value, c.index, c.mask, c.n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

typedef int (*WideFn)(const WideCall &);

template <typename T> static WideFn pick_wide(int op, bool is_int) {
    switch (op) {
        case B200_OP_ADD: return launch_scatter_wide<T, B200_OP_ADD>;
        case B200_OP_MIN: return launch_scatter_wide<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_scatter_wide<T, B200_OP_MAX>;
    }
    if constexpr (std::is_integral<T>::value) {
        if (op == B200_OP_AND)
            return launch_scatter_wide<T, B200_OP_AND>;
        if (op == B200_OP_OR)
            return launch_scatter_wide<T, B200_OP_OR>;
    }
    (void) is_int;
    return nullptr;
}

/// packet.cu (b200_scatter_reduce_idx): (vt, op) has been validated by the caller
int scatter_reduce_wide(cudaStream_t stream, int vt, int op, int mode, void *target, const void *value,
                        const uint64_t *index, const uint8_t *mask, uint64_t n) {
    WideFn fn = nullptr;
    switch (vt) {
        case B200_VT_INT32:   fn = pick_wide<int32_t>(op, true); break;
        case B200_VT_UINT32:  fn = pick_wide<uint32_t>(op, true); break;
        case B200_VT_INT64:   fn = pick_wide<long long>(op, true); break;
        case B200_VT_UINT64:  fn = pick_wide<unsigned long long>(op, true); break;
        case B200_VT_FLOAT16: fn = pick_wide<__half>(op, false); break;
        case B200_VT_FLOAT32: fn = pick_wide<float>(op, false); break;
        case B200_VT_FLOAT64: fn = pick_wide<double>(op, false); break;
    }
    if (!fn)
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_var_scatter(): the %s backend does not support the requested type of "
                    "atomic reduction (%s) for variables of type (%s)", "CUDA", op_name(op), type_name(vt));
    WideCall call{ stream, target, value, index, mask, n, mode };
    return fn(call);
}

/// packet.cu: float16 Add packets (red.global.v2 / .v4 / .v8.f16)
int packet_f16_add(cudaStream_t stream, void *target, const void *const *values, uint32_t width,
                   const uint32_t *index, const uint8_t *mask, uint64_t n, int mode);

} // namespace b200

using namespace b200;

extern "C" {

int b200_can_scatter_reduce(int vt, int op) {
    // src/op.cpp:2735-2820 evaluated for the CUDA backend at compute capability 10.0
    // (Identity -- the plain scatter -- is served by b200_scatter_reduce_idx)
    if (op == B200_OP_IDENTITY)
        return type_size(vt) != 0;
    return pick_scatter(vt, op) != nullptr;
}

int b200_scatter_reduce(void *stream_, int vt, int op, int mode, void *target,
                        const void *value, const uint32_t *index, const uint8_t *mask,
                        uint64_t n) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (op == B200_OP_IDENTITY) // plain scatter (packet.cu)
        return b200_scatter_reduce_idx(stream_, vt, op, mode, target, value, index, B200_VT_UINT32, mask, n);
    ScatterFn fn = pick_scatter(vt, op);
    if (!fn)
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_var_scatter(): the %s backend does not support the requested type of "
                    "atomic reduction (%s) for variables of type (%s)",
                    "CUDA", op_name(op), type_name(vt));
    if (mode != B200_MODE_AUTO && mode != B200_MODE_DIRECT && mode != B200_MODE_LOCAL && mode != B200_MODE_NO_CONFLICTS)
        return fail(B200_ERR_UNSUPPORTED, "jit_var_scatter(): unsupported reduction mode %d", mode);
    if (n == 0)
        return B200_OK;
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, B200_KERNEL_SCATTER, n);
    ScatterCall call{ stream, target, value, index, mask, n, mode };
    return fn(call);
}

int b200_scatter_inc(void *stream_, uint32_t *target, const uint32_t *index, const uint8_t *mask,
                     uint32_t *out, uint64_t n) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (n == 0)
        return B200_OK;
    cudaStream_t stream = resolve_stream(stream_);
    if (n >= 4096 && ((uintptr_t) index % 16) == 0 && ((uintptr_t) out % 16) == 0 &&
        (!mask || ((uintptr_t) mask % 4) == 0)) {
        // four entries per thread; the (< 4) tail entries go through the scalar kernel
        const uint64_t nvec = n / 4;
        uint32_t vgrid = (uint32_t) std::max<uint64_t>(
            1, std::min<uint64_t>(ceil_div(nvec, (uint64_t) SCATTER_THREADS), (uint64_t) sm_count() * 8));
        scatter_inc_vec4_kernel<<<vgrid, SCATTER_THREADS, 0, stream>>>(target, index, mask, out, nvec);
        B200_LAUNCH_CHECK();
        if (n % 4 == 0)
            return B200_OK;
        index += 4 * nvec;
        out += 4 * nvec;
        if (mask)
            mask += 4 * nvec;
        n %= 4;
    }
    uint32_t grid = (uint32_t) std::max<uint64_t>(
        1, std::min<uint64_t>(ceil_div(n, (uint64_t) SCATTER_THREADS), (uint64_t) sm_count() * 16));
    scatter_inc_kernel<<<grid, SCATTER_THREADS, 0, stream>>>(target, index, mask, out, n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int b200_scatter_reduce_packet(void *stream_, int vt, int op, int mode, void *target,
                               const void *const *values, uint32_t width, const uint32_t *index,
                               const uint8_t *mask, uint64_t n) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (width == 0 || (width & (width - 1)) != 0)
        return fail(B200_ERR_INVALID, "jit_var_scatter_packet(): vector size must be a power of two!");
    if (mode != B200_MODE_AUTO && mode != B200_MODE_DIRECT && mode != B200_MODE_LOCAL)
        return fail(B200_ERR_UNSUPPORTED, "jit_var_scatter_packet(): unsupported reduction mode %d", mode);
    if (vt == B200_VT_FLOAT16 && op == B200_OP_ADD) {
        if (width > 8)
            return fail(B200_ERR_UNSUPPORTED,
                        "jit_var_scatter_packet(): packet size must be 1, 2, 4 or 8 (got %u)", width);
        // red.global.v2 / .v4 / .v8.f16 need naturally aligned packets
        if ((uintptr_t) target % (width * 2) != 0)
            return fail(B200_ERR_INVALID,
                        "jit_var_scatter_packet(): the target of a float16 packet reduction must be "
                        "aligned to %u bytes!", width * 2);
        if (n == 0)
            return B200_OK;
        cudaStream_t stream = resolve_stream(stream_);
        HistoryScope hs(stream, B200_KERNEL_SCATTER, n);
        return packet_f16_add(stream, target, values, width, index, mask, n, mode);
    }
    PacketFn fn = pick_packet(vt, op);
    if (!fn) {
        // no packet kernel for this (type, op): component by component
        if (!pick_scatter(vt, op))
            return fail(B200_ERR_UNSUPPORTED,
                        "jit_var_scatter_packet(): the %s backend does not support the requested type of "
                        "atomic reduction (%s) for variables of type (%s)",
                        "CUDA", op_name(op), type_name(vt));
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_var_scatter_packet(): no packet kernel for (%s, %s); scatter the components "
                    "one by one", type_name(vt), op_name(op));
    }
    if (n == 0)
        return B200_OK;
    // red.global.add.v2 / .v4.f32 need 8- / 16-byte aligned packets (a misaligned
    // address would raise a sticky misaligned-address fault)
    if (vt == B200_VT_FLOAT32 && op == B200_OP_ADD && width >= 2 &&
        ((uintptr_t) target % (std::min<uint32_t>(width, 4) * 4)) != 0)
        return fail(B200_ERR_INVALID,
                    "jit_var_scatter_packet(): the target of a float32 packet reduction must be "
                    "aligned to %u bytes!", std::min<uint32_t>(width, 4) * 4);
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, B200_KERNEL_SCATTER, n);
    PacketCall call{ stream, target, values, width, index, mask, n, mode };
    return fn(call);
}

} // extern "C"
