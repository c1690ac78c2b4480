// reduce.cu -- block reductions, whole-array reductions and dot products.
//
// Replaces CUDAThreadState::block_reduce / reduce_dot (src/cuda_ts.cpp:195-398)
// and the kernels resources/block_reduce.cuh, resources/reduce_2.cuh.
//
// The reference launches one thread per element (or per 16-byte vector) and
// recurses through up to three launches with temporaries.  Here every block
// size is served by one of three streaming kernels, all built on 128-bit loads
// with several loads in flight per thread:
//
//   rows    block_size is a power of two and a block is at most 8 warp rows
//           (4 KiB): in-register / in-warp butterflies, no shared memory.
//   tiles   other blocks below 512 B (odd sizes, misaligned arrays): a CTA
//           stages a whole number of blocks in shared memory with aligned
//           vector loads and reduces them with sub-warp groups.
//   chunks  blocks of 512 B and more: one warp or one CTA ("team") streams a
//           contiguous chunk; blocks that are larger than a chunk produce
//           partials that a tiny finishing kernel combines in a fixed order
//           (run-to-run deterministic, also for floating point).
#include "common.cuh"

#include <atomic>

#include <cstdlib>

namespace b200 {

static constexpr int REDUCE_THREADS = 256;

// ----------------------------------------------------------------- helpers

/// Load the N elements of vector 'q' of the 16-byte-aligned virtual array
/// (vbase + q*N), converting to the value type.  Elements whose real index
/// (virtual index - mis) lies outside [lo, hi) are replaced by the identity.
template <typename T, int Op>
B200_DEVICE void load_guarded(const T *in, uint64_t q, uint32_t mis, uint64_t lo,
                              uint64_t hi, typename ValueOf<T>::type *e) {
    using V = typename ValueOf<T>::type;
    constexpr int N = VecInfo<T>::N;
    uint64_t v0 = q * N; // virtual index of element 0
    if (v0 >= lo + mis && v0 + N <= hi + mis) {
        Vec16<T> v;
        v.raw = ld_stream(in + (v0 - mis));
        #pragma unroll
        for (int k = 0; k < N; ++k)
            e[k] = to_value<T>(v.elem[k]);
    } else {
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            uint64_t vi = v0 + k;
            bool ok = vi >= lo + mis && vi < hi + mis;
            e[k] = ok ? to_value<T>(in[vi - mis]) : Red<V, Op>::identity();
        }
    }
}

// ------------------------------------------------------------------- rows

template <int BYTES> B200_DEVICE void store_packed(void *dst, const void *src) {
    if constexpr (BYTES == 1) *(uint8_t *) dst = *(const uint8_t *) src;
    else if constexpr (BYTES == 2) *(uint16_t *) dst = *(const uint16_t *) src;
    else if constexpr (BYTES == 4) *(uint32_t *) dst = *(const uint32_t *) src;
    else if constexpr (BYTES == 8) *(uint2 *) dst = *(const uint2 *) src;
    else *(uint4 *) dst = *(const uint4 *) src;
}

/// 2^L elements per block, N / 2^L blocks inside one 16-byte vector: pairwise
/// tree in registers, then one packed store of the N >> L results.
template <typename T, int Op, int L>
B200_DEVICE void rows_in_vector(typename ValueOf<T>::type *e, T *out, uint64_t q,
                                uint64_t nblocks) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr int PER_VEC = N >> L;
    #pragma unroll
    for (int s = 0; s < L; ++s) {
        #pragma unroll
        for (int i = 0; i < N; i += 2 << s)
            e[i] = R::apply(e[i], e[i + (1 << s)]);
    }
    T res[PER_VEC];
    #pragma unroll
    for (int i = 0; i < PER_VEC; ++i)
        res[i] = from_value<T>(e[i << L]);
    uint64_t o = q * PER_VEC;
    if (o + PER_VEC <= nblocks) {
        store_packed<PER_VEC * (int) sizeof(T)>(out + o, res);
    } else {
        #pragma unroll
        for (int i = 0; i < PER_VEC; ++i)
            if (o + i < nblocks)
                out[o + i] = res[i];
    }
}

/// Power-of-two blocks of at most U warp rows (U * 512 bytes), 16-byte aligned
/// input.  Every warp iteration streams U consecutive rows of 32 vectors, all
/// U loads of a lane in flight at once.  Three regimes, selected by a
/// warp-uniform branch on log2(block_size):
///   (a) several blocks inside a vector: in-register tree + packed store;
///   (b) a block spans 2..32 adjacent lanes: butterfly over those lanes;
///   (c) a block spans 2..U whole rows: in-register tree over the lane's row
///       partials, then a full-warp butterfly.
/// Full-warp reduction of C values per lane at once (C a power of two): at every
/// level the lanes with 'BIT' set keep the upper half of the values and send the
/// lower half to their partner (and vice versa), so C values cost C - 1 + log2(32 / C)
/// shuffles instead of 5 C.  Returns, in every lane, the reduction of value
/// j = lane >> (5 - log2 C) over all 32 lanes.  (A butterfly per value made the
/// 512-byte blocks the slowest block size: 40 shuffles per 4 KiB.)
template <typename V, int Op, int C, int BIT = 16>
B200_DEVICE V warp_fold(V (&r)[C], uint32_t lane) {
    using R = Red<V, Op>;
    if constexpr (C == 1) {
        V x = r[0];
        #pragma unroll
        for (int d = BIT; d > 0; d >>= 1)
            x = R::apply(x, shfl_xor(x, d));
        return x;
    } else {
        constexpr int H = C / 2;
        V a[H];
        const bool up = (lane & BIT) != 0;
        #pragma unroll
        for (int i = 0; i < H; ++i) {
            const V send = up ? r[i] : r[i + H];
            const V keep = up ? r[i + H] : r[i];
            a[i] = R::apply(keep, shfl_xor(send, BIT));
        }
        return warp_fold<V, Op, H, BIT / 2>(a, lane);
    }
}

template <typename T, int Op, int U>
__global__ void __launch_bounds__(REDUCE_THREADS, sizeof(T) >= 4 ? 4 : 2)
reduce_rows_kernel(const T *__restrict__ in, T *__restrict__ out, uint64_t size,
                   uint32_t log2_bs) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr int LOG2_N = N == 2 ? 1 : N == 4 ? 2 : N == 8 ? 3 : 4;
    constexpr int LOG2_U = U == 2 ? 1 : U == 4 ? 2 : U == 8 ? 3 : 4;

    const uint64_t nvec = (size + N - 1) / N;
    const uint64_t nrows = (nvec + 31) / 32;
    const uint64_t nblocks = (size + (1ull << log2_bs) - 1) >> log2_bs;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t) blockIdx.x * REDUCE_THREADS + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t) gridDim.x * REDUCE_THREADS) >> 5;

    for (uint64_t row0 = warp * U; row0 < nrows; row0 += nwarps * U) {
        V e[U][N];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t q = (row0 + u) * 32 + lane;
            if (q < nvec) {
                load_guarded<T, Op>(in, q, 0, 0, size, e[u]);
            } else {
                #pragma unroll
                for (int k = 0; k < N; ++k)
                    e[u][k] = R::identity();
            }
        }

        if (log2_bs <= LOG2_N) { // (a)
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                uint64_t q = (row0 + u) * 32 + lane;
                if (q < nvec) {
                    switch (log2_bs) {
                        case 1: rows_in_vector<T, Op, 1>(e[u], out, q, nblocks); break;
                        case 2: if constexpr (LOG2_N >= 2) rows_in_vector<T, Op, 2>(e[u], out, q, nblocks); break;
                        case 3: if constexpr (LOG2_N >= 3) rows_in_vector<T, Op, 3>(e[u], out, q, nblocks); break;
                        case 4: if constexpr (LOG2_N >= 4) rows_in_vector<T, Op, 4>(e[u], out, q, nblocks); break;
                    }
                }
            }
            continue;
        }

        // reduce every vector to a single value
        V r[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            r[u] = e[u][0];
            #pragma unroll
            for (int k = 1; k < N; ++k)
                r[u] = R::apply(r[u], e[u][k]);
        }

        if (log2_bs <= LOG2_N + 5) { // (b)
            const uint32_t shift = log2_bs - LOG2_N; // log2(lanes per block), 1..5
            const int lanes = 1 << shift;
            if (shift >= 3 && U == 8) {
                // 8 .. 32 lanes per block: the U rows are reduced together, folding over
                // the top three lane bits of a block; lane gets row (lane >> (shift - 3)) & 7
                V x;
                if (shift == 5)
                    x = warp_fold<V, Op, U, 16>(r, lane);
                else if (shift == 4)
                    x = warp_fold<V, Op, U, 8>(r, lane);
                else
                    x = warp_fold<V, Op, U, 4>(r, lane);
                const uint32_t j = (lane >> (shift - 3)) & 7u;
                const uint64_t o = ((row0 + j) << (5 - shift)) + (lane >> shift);
                if ((lane & ((1u << (shift - 3)) - 1u)) == 0 && o < nblocks)
                    out[o] = from_value<T>(x);
                continue;
            }
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                uint64_t q = (row0 + u) * 32 + lane;
                V x = warp_reduce<V, Op>(r[u], lanes);
                if ((lane & (lanes - 1)) == 0 && (q >> shift) < nblocks)
                    out[q >> shift] = from_value<T>(x);
            }
        } else { // (c)
            const uint32_t log2_rows = log2_bs - LOG2_N - 5; // 1..LOG2_U
            #pragma unroll
            for (int s = 0; s < LOG2_U; ++s) {
                if ((int) log2_rows > s) {
                    #pragma unroll
                    for (int u = 0; u < U; u += 2 << s)
                        r[u] = R::apply(r[u], r[u + (1 << s)]);
                }
            }
            // the U >> log2_rows remaining values are reduced over the warp together
            auto finish = [&](auto tag) {
                constexpr int C = decltype(tag)::value;
                constexpr int LOG2_C = C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : 3;
                V v[C];
                #pragma unroll
                for (int i = 0; i < C; ++i)
                    v[i] = r[i * (U / C)];
                const V x = warp_fold<V, Op, C>(v, lane);
                const uint64_t o = (row0 >> log2_rows) + (lane >> (5 - LOG2_C));
                if ((lane & ((32 >> LOG2_C) - 1)) == 0 && o < nblocks)
                    out[o] = from_value<T>(x);
            };
            switch (U >> log2_rows) {
                case 1: finish(std::integral_constant<int, 1>()); break;
                case 2: if constexpr (U >= 2) finish(std::integral_constant<int, 2>()); break;
                case 4: if constexpr (U >= 4) finish(std::integral_constant<int, 4>()); break;
                default: break; // log2_rows >= 1: at most U / 2 values
            }
        }
    }
}

// ------------------------------------------------------------------ tiles

/// Arbitrary small blocks (block bytes < 512): each CTA owns 'nb' whole blocks,
/// staged in shared memory.  Dynamic shared memory: (cap + 2 * N) elements.
/// The blocks are then reduced by groups of 2^log2_group lanes each (one lane per block
/// when log2_group == 0: consecutive lanes read shared memory bs elements apart, which is
/// conflict free for the block sizes the launcher sends here).  All indices inside a tile
/// are 32 bit; tiles that lie inside the array take loads without bounds checks.
template <typename T, int Op>
__global__ void __launch_bounds__(REDUCE_THREADS)
reduce_tiles_kernel(const T *__restrict__ in, T *__restrict__ out, uint64_t size,
                    uint32_t bs, uint32_t nb, uint32_t log2_group) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    extern __shared__ uint4 tile_smem[];
    T *tile = (T *) tile_smem;

    const uint32_t mis = (uint32_t) (((uintptr_t) in & 15) / sizeof(T));
    const uint64_t nblocks = (size + bs - 1) / bs;
    const uint64_t blk0 = (uint64_t) blockIdx.x * nb;
    const uint64_t start = blk0 * bs;
    const uint64_t end = min(start + (uint64_t) nb * bs, size);
    const uint32_t nb_tile = (uint32_t) min((uint64_t) nb, nblocks - blk0);

    const uint64_t q_first = (start + mis) / N, q_last = (end - 1 + mis) / N;
    const uint32_t nvec = (uint32_t) (q_last - q_first + 1);
    // offset of element 'start' inside the staged tile
    const uint32_t shift = (uint32_t) (start + mis - q_first * N);

    if (q_first * N >= mis && (q_last + 1) * N <= size + mis) {
        // (vectors may straddle into a neighbour tile)
        const uint4 *src = (const uint4 *) (in + (q_first * N - mis));
        for (uint32_t i = threadIdx.x; i < nvec; i += REDUCE_THREADS)
            tile_smem[i] = ld_stream(src + i);
    } else {
        for (uint32_t i = threadIdx.x; i < nvec; i += REDUCE_THREADS) {
            uint64_t q = q_first + i, v0 = q * N;
            Vec16<T> v;
            if (v0 >= mis && v0 + N <= size + mis) {
                v.raw = ld_stream(in + (v0 - mis));
            } else {
                #pragma unroll
                for (int k = 0; k < N; ++k) {
                    uint64_t vi = v0 + k;
                    v.elem[k] = (vi >= mis && vi < size + mis) ? in[vi - mis] : T();
                }
            }
            tile_smem[i] = v.raw;
        }
    }
    __syncthreads();

    const uint32_t tail = (uint32_t) (end - start) - (nb_tile - 1) * bs; // length of the tile's last block
    out += blk0;
    if (log2_group == 0) {
        // one thread per block
        for (uint32_t b = threadIdx.x; b < nb_tile; b += REDUCE_THREADS) {
            const T *src = tile + shift + b * bs;
            const uint32_t len = b + 1 < nb_tile ? bs : tail;
            V acc = to_value<T>(src[0]);
            #pragma unroll 4
            for (uint32_t k = 1; k < len; ++k)
                acc = R::apply(acc, to_value<T>(src[k]));
            out[b] = from_value<T>(acc);
        }
        return;
    }

    // groups of G lanes per block (G a compile-time constant: the loops unroll)
    auto by_groups = [&](auto tag) {
        constexpr uint32_t LOG2_G = decltype(tag)::value, G = 1u << LOG2_G, GROUPS_PER_WARP = 32 >> LOG2_G;
        constexpr uint32_t WARPS = REDUCE_THREADS / 32;
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t sub = lane & (G - 1), grp = lane >> LOG2_G;
        // BU blocks per group and step: their shuffle reductions are independent chains
        constexpr uint32_t BU = 4;
        for (uint32_t b0 = warp * GROUPS_PER_WARP; b0 < nb_tile; b0 += WARPS * GROUPS_PER_WARP * BU) {
            V acc[BU];
            #pragma unroll
            for (uint32_t u = 0; u < BU; ++u) {
                const uint32_t b = b0 + u * WARPS * GROUPS_PER_WARP + grp;
                acc[u] = R::identity();
                if (b < nb_tile) {
                    const uint32_t len = b + 1 < nb_tile ? bs : tail;
                    const T *src = tile + shift + b * bs;
                    // (a block has fewer than 2 * G elements, except for G == 32: fewer than 512 bytes)
                    constexpr uint32_t ROUNDS = G == 32 ? (uint32_t) ((512 / sizeof(T) + 31) / 32) : 2u;
                    #pragma unroll
                    for (uint32_t r = 0; r < ROUNDS; ++r) {
                        const uint32_t k = sub + r * G;
                        if (k < len)
                            acc[u] = R::apply(acc[u], to_value<T>(src[k]));
                    }
                }
            }
            #pragma unroll
            for (uint32_t d = G >> 1; d > 0; d >>= 1) {
                #pragma unroll
                for (uint32_t u = 0; u < BU; ++u)
                    acc[u] = R::apply(acc[u], shfl_xor(acc[u], (int) d));
            }
            #pragma unroll
            for (uint32_t u = 0; u < BU; ++u) {
                const uint32_t b = b0 + u * WARPS * GROUPS_PER_WARP + grp;
                if (b < nb_tile && sub == 0)
                    out[b] = from_value<T>(acc[u]);
            }
        }
    };
    switch (log2_group) {
        case 1: by_groups(std::integral_constant<uint32_t, 1>()); break;
        case 2: by_groups(std::integral_constant<uint32_t, 2>()); break;
        case 3: by_groups(std::integral_constant<uint32_t, 3>()); break;
        case 4: by_groups(std::integral_constant<uint32_t, 4>()); break;
        default: by_groups(std::integral_constant<uint32_t, 5>()); break;
    }
}

// ----------------------------------------------------------------- chunks

/// One team (TEAM = 32: a warp, TEAM = 256: the CTA) reduces one chunk of one
/// block: [b * bs + c * chunk, min(+chunk, block end, size)).  Results go to
/// out[b * chunks_per_block + c] in the VALUE type when PARTIAL, else as T.
template <typename T, int Op, int TEAM, bool PARTIAL, int U>
__global__ void __launch_bounds__(REDUCE_THREADS)
reduce_chunks_kernel(const T *__restrict__ in, void *__restrict__ out_, uint64_t size,
                     uint64_t bs, uint64_t chunk, uint32_t chunks_per_block,
                     uint64_t nteams, unsigned int *ticket = nullptr, T *final_out = nullptr,
                     uint64_t nblocks = 0) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr int TEAMS_PER_CTA = REDUCE_THREADS / TEAM;
    __shared__ V warp_part[REDUCE_THREADS / 32];

    const uint32_t mis = (uint32_t) (((uintptr_t) in & 15) / sizeof(T));
    const uint32_t t = threadIdx.x % TEAM;

    for (uint64_t p = (uint64_t) blockIdx.x * TEAMS_PER_CTA + threadIdx.x / TEAM;
         p < nteams; p += (uint64_t) gridDim.x * TEAMS_PER_CTA) {
        uint64_t b = p / chunks_per_block, c = p - b * chunks_per_block;
        uint64_t start = b * bs + c * chunk;
        uint64_t end = min(min(start + chunk, (b + 1) * bs), size);

        V acc[U];
        #pragma unroll
        for (int u = 0; u < U; ++u)
            acc[u] = R::identity();

        if (start < end) {
            uint64_t q_first = (start + mis) / N, q_last = (end - 1 + mis) / N;
            // guarded boundary vectors
            if (t == 0) {
                V e[N];
                load_guarded<T, Op>(in, q_first, mis, start, end, e);
                #pragma unroll
                for (int k = 0; k < N; ++k)
                    acc[0] = R::apply(acc[0], e[k]);
            }
            if (t == TEAM - 1 && q_last != q_first) {
                V e[N];
                load_guarded<T, Op>(in, q_last, mis, start, end, e);
                #pragma unroll
                for (int k = 0; k < N; ++k)
                    acc[U - 1] = R::apply(acc[U - 1], e[k]);
            }
            // interior vectors are entirely inside [start, end)
            const uint4 *vbase = (const uint4 *) (in - mis);
            uint64_t q = q_first + 1 + t;
            for (; q + (uint64_t) (U - 1) * TEAM < q_last; q += (uint64_t) U * TEAM) {
                Vec16<T> v[U];
                #pragma unroll
                for (int u = 0; u < U; ++u)
                    v[u].raw = ld_stream(vbase + q + (uint64_t) u * TEAM);
                #pragma unroll
                for (int u = 0; u < U; ++u) {
                    #pragma unroll
                    for (int k = 0; k < N; ++k)
                        acc[u] = R::apply(acc[u], to_value<T>(v[u].elem[k]));
                }
            }
            for (; q < q_last; q += TEAM) {
                Vec16<T> v;
                v.raw = ld_stream(vbase + q);
                #pragma unroll
                for (int k = 0; k < N; ++k)
                    acc[0] = R::apply(acc[0], to_value<T>(v.elem[k]));
            }
        }

        V r = acc[0];
        #pragma unroll
        for (int u = 1; u < U; ++u)
            r = R::apply(r, acc[u]);
        r = warp_reduce<V, Op>(r);

        if constexpr (TEAM > 32) {
            __syncthreads(); // warp_part reuse across iterations
            if ((threadIdx.x & 31) == 0)
                warp_part[threadIdx.x >> 5] = r;
            __syncthreads();
            if (threadIdx.x < 32) {
                r = threadIdx.x < REDUCE_THREADS / 32 ? warp_part[threadIdx.x]
                                                       : R::identity();
                r = warp_reduce<V, Op>(r);
            }
        }

        if (t == 0) {
            if constexpr (PARTIAL)
                ((V *) out_)[p] = r;
            else
                ((T *) out_)[p] = from_value<T>(r);
        }
    }

    // PARTIAL with a ticket: the CTA that finishes last combines the partials of every
    // block in a fixed order (a warp per block, as reduce_finish_kernel) -- one launch,
    // run-to-run deterministic floating point sums -- and leaves the ticket at zero
    if constexpr (PARTIAL) {
        if (ticket) {
            __shared__ bool last;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0)
                last = atomicAdd(ticket, 1u) == gridDim.x - 1;
            __syncthreads();
            if (!last)
                return;
            __threadfence();
            const V *partials = (const V *) out_;
            const uint32_t lane = threadIdx.x & 31;
            if (nblocks == 1) {
                // the whole CTA on one block: thread t takes the partials t, t + 256, ... (four
                // independent loads in flight), then the usual warp / CTA tree
                V a4[4] = { R::identity(), R::identity(), R::identity(), R::identity() };
                uint32_t i = threadIdx.x;
                for (; i + 3 * REDUCE_THREADS < chunks_per_block; i += 4 * REDUCE_THREADS) {
                    V v[4];
                    #pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = __ldcg(partials + i + u * REDUCE_THREADS);
                    #pragma unroll
                    for (int u = 0; u < 4; ++u)
                        a4[u] = R::apply(a4[u], v[u]);
                }
                for (; i < chunks_per_block; i += REDUCE_THREADS)
                    a4[0] = R::apply(a4[0], __ldcg(partials + i));
                V acc = warp_reduce<V, Op>(R::apply(R::apply(a4[0], a4[1]), R::apply(a4[2], a4[3])));
                __syncthreads(); // warp_part reuse
                if (lane == 0)
                    warp_part[threadIdx.x >> 5] = acc;
                __syncthreads();
                if (threadIdx.x < 32) {
                    acc = threadIdx.x < REDUCE_THREADS / 32 ? warp_part[threadIdx.x] : R::identity();
                    acc = warp_reduce<V, Op>(acc);
                    if (threadIdx.x == 0)
                        final_out[0] = from_value<T>(acc);
                }
            } else {
                for (uint64_t b = threadIdx.x >> 5; b < nblocks; b += REDUCE_THREADS / 32) {
                    V acc = R::identity();
                    for (uint32_t i = lane; i < chunks_per_block; i += 32)
                        acc = R::apply(acc, __ldcg(partials + b * chunks_per_block + i));
                    acc = warp_reduce<V, Op>(acc);
                    if (lane == 0)
                        final_out[b] = from_value<T>(acc);
                }
            }
            if (threadIdx.x == 0)
                *ticket = 0;
        }
    }
}

/// Combine 'count' value-typed partials per block in a fixed order (one warp
/// per block) and narrow to T.
template <typename T, int Op>
__global__ void __launch_bounds__(REDUCE_THREADS)
reduce_finish_kernel(const typename ValueOf<T>::type *__restrict__ partials,
                     T *__restrict__ out, uint32_t count, uint64_t nblocks) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    uint32_t lane = threadIdx.x & 31;
    for (uint64_t b = (uint64_t) blockIdx.x * (REDUCE_THREADS / 32) + (threadIdx.x >> 5);
         b < nblocks; b += (uint64_t) gridDim.x * (REDUCE_THREADS / 32)) {
        V acc = R::identity();
        for (uint32_t i = lane; i < count; i += 32)
            acc = R::apply(acc, partials[b * count + i]);
        acc = warp_reduce<V, Op>(acc);
        if (lane == 0)
            out[b] = from_value<T>(acc);
    }
}

// -------------------------------------------------------------------- dot

/// Per-CTA partial of sum a[i] * b[i] (fused multiply-add in the value type).
template <typename T, int U>
__global__ void __launch_bounds__(REDUCE_THREADS)
dot_chunks_kernel(const T *__restrict__ a, const T *__restrict__ b,
                  typename ValueOf<T>::type *__restrict__ partials, uint64_t size,
                  uint64_t chunk, bool aligned) {
    using V = typename ValueOf<T>::type;
    constexpr int N = VecInfo<T>::N;
    __shared__ V warp_part[REDUCE_THREADS / 32];

    uint64_t start = (uint64_t) blockIdx.x * chunk, end = min(start + chunk, size);
    V acc[U];
    #pragma unroll
    for (int u = 0; u < U; ++u)
        acc[u] = (V) 0;

    if (aligned) { // chunk is a multiple of N and both arrays are 16-byte aligned
        uint64_t nfull = (end - start) / N;
        const uint4 *va = (const uint4 *) (a + start), *vb = (const uint4 *) (b + start);
        uint64_t q = threadIdx.x;
        for (; q + (uint64_t) (U - 1) * REDUCE_THREADS < nfull; q += (uint64_t) U * REDUCE_THREADS) {
            Vec16<T> x[U], y[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                x[u].raw = ld_stream(va + q + u * REDUCE_THREADS);
                y[u].raw = ld_stream(vb + q + u * REDUCE_THREADS);
            }
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                #pragma unroll
                for (int k = 0; k < N; ++k)
                    acc[u] = fma(to_value<T>(x[u].elem[k]), to_value<T>(y[u].elem[k]), acc[u]);
            }
        }
        for (; q < nfull; q += REDUCE_THREADS) {
            Vec16<T> x, y;
            x.raw = ld_stream(va + q);
            y.raw = ld_stream(vb + q);
            #pragma unroll
            for (int k = 0; k < N; ++k)
                acc[0] = fma(to_value<T>(x.elem[k]), to_value<T>(y.elem[k]), acc[0]);
        }
        for (uint64_t i = start + nfull * N + threadIdx.x; i < end; i += REDUCE_THREADS)
            acc[0] = fma(to_value<T>(a[i]), to_value<T>(b[i]), acc[0]);
    } else {
        for (uint64_t i = start + threadIdx.x; i < end; i += REDUCE_THREADS)
            acc[0] = fma(to_value<T>(a[i]), to_value<T>(b[i]), acc[0]);
    }

    V r = acc[0];
    #pragma unroll
    for (int u = 1; u < U; ++u)
        r += acc[u];
    r = warp_reduce<V, B200_OP_ADD>(r);
    if ((threadIdx.x & 31) == 0)
        warp_part[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x < 32) {
        r = threadIdx.x < REDUCE_THREADS / 32 ? warp_part[threadIdx.x] : (V) 0;
        r = warp_reduce<V, B200_OP_ADD>(r);
        if (threadIdx.x == 0)
            partials[blockIdx.x] = r;
    }
}

// ---------------------------------------------------------------- dispatch

struct ReduceCall {
    cudaStream_t stream;
    const void *in;
    void *out;
    uint64_t size, bs;
};

template <typename T, int Op> static int launch_block_reduce(const ReduceCall &c) {
    using V = typename ValueOf<T>::type;
    constexpr int N = VecInfo<T>::N;
    const T *in = (const T *) c.in;
    T *out = (T *) c.out;
    const uint64_t size = c.size, bs = c.bs;
    const uint64_t nblocks = ceil_div(size, bs);
    const uint64_t block_bytes = bs * sizeof(T);
    const int sms = sm_count();

    bool in_aligned = ((uintptr_t) in % 16) == 0;

    constexpr int ROWS_U = 8;
    if (block_bytes <= 512 * ROWS_U && is_pow2(bs) && in_aligned) {
        uint32_t log2_bs = log2i(bs);
        uint32_t store_bytes = bs <= (uint64_t) N ? (uint32_t) (N / bs * sizeof(T)) : (uint32_t) sizeof(T);
        if (((uintptr_t) out % store_bytes) == 0) {
            uint64_t nrows = ceil_div(ceil_div(size, N), 32);
            uint32_t grid = (uint32_t) std::min<uint64_t>(
                ceil_div(nrows, (uint64_t) (REDUCE_THREADS / 32) * ROWS_U), (uint64_t) sms * 16);
            reduce_rows_kernel<T, Op, ROWS_U><<<grid, REDUCE_THREADS, 0, c.stream>>>(in, out, size, log2_bs);
            B200_LAUNCH_CHECK();
            return B200_OK;
        }
    }

    if (block_bytes < 512) {
        constexpr uint32_t CAP = 32768 / sizeof(T);
        uint32_t nb = (uint32_t) (CAP / bs);
        uint32_t group = 1, log2_group = 0;
        while (group * 2 <= bs && group < 32) { group *= 2; log2_group++; }
        // short blocks whose lanes would not collide in shared memory (consecutive blocks start
        // an odd number of 32-bit words apart, or two words for 6- and 10-element blocks of
        // 4-byte types ...): one thread per block
        {
            const uint32_t words2 = (uint32_t) (bs * sizeof(T) / 2); // block stride in half words
            const uint32_t conflict = words2 % 2 ? 1u : std::min(32u, (words2 / 2) & (~(words2 / 2) + 1u));
            if (bs <= 32 && conflict <= 2)
                group = 1, log2_group = 0;
        }
        uint64_t ntiles = ceil_div(nblocks, nb);
        if (ntiles > 0x7fffffffull)
            return fail(B200_ERR_INVALID, "jit_block_reduce(): array too large!");
        size_t smem = (size_t) (CAP + 2 * N) * sizeof(T);
        reduce_tiles_kernel<T, Op><<<(uint32_t) ntiles, REDUCE_THREADS, smem, c.stream>>>(
            in, out, size, (uint32_t) bs, nb, log2_group);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }

    // chunked streaming: pick the team size and the number of chunks per block
    constexpr int U = 8;
    const uint64_t iter_bytes_cta = (uint64_t) REDUCE_THREADS * U * 16; // 32 KiB
    const uint64_t target_teams = (uint64_t) sms * 8;
    if (block_bytes <= 65536 && nblocks >= (uint64_t) sms * 32) {
        // a warp per block
        uint32_t grid = (uint32_t) std::min<uint64_t>(ceil_div(nblocks, REDUCE_THREADS / 32),
                                                       (uint64_t) sms * 16);
        reduce_chunks_kernel<T, Op, 32, false, 4><<<grid, REDUCE_THREADS, 0, c.stream>>>(
            in, out, size, bs, bs, 1, nblocks);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }

    uint64_t max_chunks = std::max<uint64_t>(1, block_bytes / iter_bytes_cta);
    uint64_t chunks = std::min(max_chunks, std::max<uint64_t>(1, ceil_div(target_teams, nblocks)));
    // Many large blocks (the finish runs as a second kernel anyway): chunks of ~128 KiB, i.e. several
    // teams per CTA.  With `target_teams` alone 256 blocks of 4 MiB became 1280 teams of 0.8 MiB on 1184
    // resident CTAs -- a second wave of 96 CTAs that took as long as the first (2^28 fp32, blocks of 2^20 /
    // 3 * 2^20 elements: 0.215 / 0.232 ms -> 0.184 / 0.186 ms).
    static const int tune = getenv("B200_REDUCE_TUNE") ? atoi(getenv("B200_REDUCE_TUNE")) : 7; // (development switch: bit 0 / 1 the two rules below, bit 2 four teams per CTA)
    if (nblocks > 64 && block_bytes >= (1u << 20) && (tune & 1))
        chunks = std::min(max_chunks, std::max(chunks, ceil_div(block_bytes, (uint64_t) 128 * 1024)));
    // Few blocks (one launch, the last CTA finishes): four teams per RESIDENT CTA, never more CTAs than
    // that -- the partial kernel holds 5 CTAs per SM (44 registers), so the 8 teams per SM assumed above
    // ran as 1.6 waves (whole-array reduce, 2^28 fp32: 0.184 -> 0.178 ms)
    uint32_t resident = 0;
    if (nblocks <= 64 && (tune & 2)) {
        static std::atomic<int> occ_cache{0};
        int occ = occ_cache.load(std::memory_order_relaxed);
        if (occ == 0) {
            B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &occ, reduce_chunks_kernel<T, Op, REDUCE_THREADS, true, U>, REDUCE_THREADS, 0));
            occ = std::max(occ, 1);
            occ_cache.store(occ, std::memory_order_relaxed);
        }
        resident = (uint32_t) (sms * occ);
        chunks = std::min(max_chunks, std::max<uint64_t>(1, (uint64_t) ((tune & 4) ? 4 : (tune & 8) ? 1 : 2) * resident / nblocks));
    }
    uint64_t chunk = ceil_div(bs, chunks);
    // keep chunk boundaries on 16-byte multiples relative to the block start
    chunk = ceil_div(chunk, N * 32) * (N * 32);
    chunks = ceil_div(bs, chunk);
    uint64_t nteams = nblocks * chunks;
    uint32_t grid = (uint32_t) std::min<uint64_t>(nteams, resident ? (uint64_t) resident : (uint64_t) sms * 16);

    if (chunks == 1) {
        reduce_chunks_kernel<T, Op, REDUCE_THREADS, false, U><<<grid, REDUCE_THREADS, 0, c.stream>>>(
            in, out, size, bs, chunk, 1, nteams);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }

    V *partials = (V *) temp_alloc(nteams * sizeof(V), c.stream);
    if (!partials)
        return fail(B200_ERR_CUDA, "jit_block_reduce(): out of memory (%zu bytes)",
                    (size_t) (nteams * sizeof(V)));
    // few blocks: one launch, the last CTA finishes (no second kernel in the stream)
    static const bool two_launches = getenv("B200_REDUCE_TWO_LAUNCHES") != nullptr; // (development switch)
    unsigned int *ticket = nblocks <= 64 && !two_launches ? stream_ticket(c.stream) : nullptr;
    if (ticket) {
        reduce_chunks_kernel<T, Op, REDUCE_THREADS, true, U><<<grid, REDUCE_THREADS, 0, c.stream>>>(
            in, partials, size, bs, chunk, (uint32_t) chunks, nteams, ticket, out, nblocks);
        temp_free(partials, c.stream);
        B200_LAUNCH_CHECK();
        return B200_OK;
    }
    reduce_chunks_kernel<T, Op, REDUCE_THREADS, true, U><<<grid, REDUCE_THREADS, 0, c.stream>>>(
        in, partials, size, bs, chunk, (uint32_t) chunks, nteams);
    count_launch();
    uint32_t grid2 = (uint32_t) std::min<uint64_t>(ceil_div(nblocks, REDUCE_THREADS / 32),
                                                    (uint64_t) sms * 8);
    reduce_finish_kernel<T, Op><<<grid2, REDUCE_THREADS, 0, c.stream>>>(
        partials, out, (uint32_t) chunks, nblocks);
    temp_free(partials, c.stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

typedef int (*ReduceFn)(const ReduceCall &);

template <typename T, bool Bits> static ReduceFn pick_op(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_block_reduce<T, B200_OP_ADD>;
        case B200_OP_MUL: return launch_block_reduce<T, B200_OP_MUL>;
        case B200_OP_MIN: return launch_block_reduce<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_block_reduce<T, B200_OP_MAX>;
        case B200_OP_AND: if constexpr (Bits) return launch_block_reduce<T, B200_OP_AND>; else return nullptr;
        case B200_OP_OR:  if constexpr (Bits) return launch_block_reduce<T, B200_OP_OR>; else return nullptr;
        default: return nullptr;
    }
}

template <typename T> static ReduceFn pick_minmax(int op) {
    switch (op) {
        case B200_OP_MIN: return launch_block_reduce<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_block_reduce<T, B200_OP_MAX>;
        default: return nullptr;
    }
}

static ReduceFn pick_reduce(int vt, int op) {
    bool minmax = op == B200_OP_MIN || op == B200_OP_MAX;
    switch (vt) {
        // 8-bit: the reference only ships and/or (block_reduce.cuh:270-271)
        case B200_VT_BOOL:
        case B200_VT_INT8:
        case B200_VT_UINT8:
            if (op == B200_OP_AND) return launch_block_reduce<uint8_t, B200_OP_AND>;
            if (op == B200_OP_OR) return launch_block_reduce<uint8_t, B200_OP_OR>;
            return nullptr;
        // signed add / mul / and / or run on the unsigned kernel (cuda_ts.cpp:215-225)
        case B200_VT_INT32:  return minmax ? pick_minmax<int32_t>(op) : pick_op<uint32_t, true>(op);
        case B200_VT_UINT32: return pick_op<uint32_t, true>(op);
        case B200_VT_INT64:  return minmax ? pick_minmax<int64_t>(op) : pick_op<uint64_t, true>(op);
        case B200_VT_UINT64: return pick_op<uint64_t, true>(op);
        case B200_VT_FLOAT16: return pick_op<__half, false>(op);
        case B200_VT_FLOAT32: return pick_op<float, false>(op);
        case B200_VT_FLOAT64: return pick_op<double, false>(op);
        default: return nullptr;
    }
}

template <typename T> static int launch_dot(cudaStream_t stream, const void *a_, const void *b_,
                                            uint64_t size, void *out) {
    using V = typename ValueOf<T>::type;
    constexpr int N = VecInfo<T>::N;
    constexpr int U = 4;
    const T *a = (const T *) a_, *b = (const T *) b_;
    const int sms = sm_count();
    bool aligned = ((uintptr_t) a % 16) == 0 && ((uintptr_t) b % 16) == 0;
    uint64_t iter = (uint64_t) REDUCE_THREADS * U * N;
    uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) sms * 8, size / iter));
    uint64_t chunk = ceil_div(ceil_div(size, chunks), N * 32) * (N * 32);
    chunks = ceil_div(size, chunk);

    V *partials = (V *) temp_alloc(chunks * sizeof(V), stream);
    if (!partials)
        return fail(B200_ERR_CUDA, "jit_reduce_dot(): out of memory");
    dot_chunks_kernel<T, U><<<(uint32_t) chunks, REDUCE_THREADS, 0, stream>>>(a, b, partials, size, chunk, aligned);
    count_launch();
    reduce_finish_kernel<T, B200_OP_ADD><<<1, REDUCE_THREADS, 0, stream>>>(partials, (T *) out, (uint32_t) chunks, 1);
    temp_free(partials, stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_block_reduce(void *stream_, int vt, int op, uint64_t size, uint64_t block_size,
                      const void *in, void *out) {
    int rc = ensure_init();
    if (rc)
        return rc;
    // src/cuda_ts.cpp:201-213
    if (size == 0)
        return B200_OK;
    if (block_size == 0 || block_size > size)
        return fail(B200_ERR_INVALID,
                    "jit_block_reduce(): invalid block size (size=%llu, block_size=%llu)!",
                    (unsigned long long) size, (unsigned long long) block_size);
    uint32_t tsize = type_size(vt);
    ReduceFn fn = pick_reduce(vt, op);
    if (!fn || tsize == 0)
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_block_reduce(): no existing kernel for type=%s, op=%s!",
                    type_name(vt), op_name(op));
    if (((uintptr_t) in % tsize) != 0 || ((uintptr_t) out % tsize) != 0)
        return fail(B200_ERR_INVALID, "jit_block_reduce(): misaligned pointer!");
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, B200_KERNEL_BLOCK_REDUCE, size);
    if (block_size == 1) {
        B200_CUDA_CHECK(cudaMemcpyAsync(out, in, size * tsize, cudaMemcpyDeviceToDevice, stream));
        return B200_OK;
    }
    ReduceCall call{ stream, in, out, size, block_size };
    return fn(call);
}

int b200_reduce(void *stream, int vt, int op, const void *in, uint64_t size, void *out) {
    // src/util.cpp:42-45
    return b200_block_reduce(stream, vt, op, size, size, in, out);
}

int b200_reduce_dot(void *stream_, int vt, const void *a, const void *b, uint64_t size,
                    void *out) {
    int rc = ensure_init();
    if (rc)
        return rc;
    cudaStream_t stream = resolve_stream(stream_);
    uint32_t tsize = type_size(vt);
    if (vt != B200_VT_FLOAT16 && vt != B200_VT_FLOAT32 && vt != B200_VT_FLOAT64)
        return fail(B200_ERR_UNSUPPORTED, "jit_reduce_dot(): no existing kernel for type=%s!",
                    type_name(vt));
    if (size == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(out, 0, tsize, stream));
        return B200_OK;
    }
    HistoryScope hs(stream, B200_KERNEL_DOT, size);
    switch (vt) {
        case B200_VT_FLOAT16: return launch_dot<__half>(stream, a, b, size, out);
        case B200_VT_FLOAT32: return launch_dot<float>(stream, a, b, size, out);
        default: return launch_dot<double>(stream, a, b, size, out);
    }
}

} // extern "C"

// ---- all / any ---------------------------------------------------------------
// jitc_all / jitc_any (src/util.cpp:153-211) reduce ceil(size / 4) u32 words with
// And / Or after overwriting up to 3 bytes past the array with the neutral value
// (src/init.cpp:919-939).  Here the byte array is reduced directly (the chunk
// kernel handles any alignment and length), so nothing outside values[0, size)
// is touched.

static int bool_reduce_async(void *stream_, const uint8_t *values, uint64_t size,
                             uint8_t *out, int op) {
    int rc = b200::ensure_init();
    if (rc)
        return rc;
    cudaStream_t stream = b200::resolve_stream(stream_);
    if (size == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(out, op == B200_OP_AND ? 1 : 0, 1, stream));
        return B200_OK;
    }
    return b200_block_reduce(stream, B200_VT_UINT8, op, size, size, values, out);
}

static int bool_reduce_sync(void *stream_, const uint8_t *values, uint64_t size, int *result,
                            int op) {
    int rc = b200::ensure_init();
    if (rc)
        return rc;
    if ((rc = b200::sync_forbidden())) // jitc_all / jitc_any wait for the stream (src/util.cpp:153-180)
        return rc;
    cudaStream_t stream = b200::resolve_stream(stream_);
    uint8_t *tmp = (uint8_t *) b200::temp_alloc(4, stream);
    if (!tmp)
        return b200::fail(B200_ERR_CUDA, "jit_all/any(): out of memory");
    rc = bool_reduce_async(stream, values, size, tmp, op);
    uint8_t host = 0;
    cudaError_t err = cudaSuccess;
    if (!rc)
        err = cudaMemcpyAsync(&host, tmp, 1, cudaMemcpyDeviceToHost, stream);
    b200::temp_free(tmp, stream);
    if (rc)
        return rc;
    if (err != cudaSuccess)
        return b200::cuda_fail(err, "cudaMemcpyAsync");
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    *result = host != 0;
    return B200_OK;
}

extern "C" {

int b200_all_async(void *stream, const uint8_t *values, uint64_t size, uint8_t *out) {
    return bool_reduce_async(stream, values, size, out, B200_OP_AND);
}

int b200_any_async(void *stream, const uint8_t *values, uint64_t size, uint8_t *out) {
    return bool_reduce_async(stream, values, size, out, B200_OP_OR);
}

int b200_all(void *stream, const uint8_t *values, uint64_t size, int *result) {
    return bool_reduce_sync(stream, values, size, result, B200_OP_AND);
}

int b200_any(void *stream, const uint8_t *values, uint64_t size, int *result) {
    return bool_reduce_sync(stream, values, size, result, B200_OP_OR);
}

} // extern "C"
