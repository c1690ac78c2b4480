// scan.cu -- blocked prefix reductions (inclusive / exclusive, forward / reverse).
//
// Replaces CUDAThreadState::block_prefix_reduce (src/cuda_ts.cpp:530-681) and
// resources/block_prefix_reduce.cuh.  The reference scans one element per
// thread with a Hillis-Steele pass through shared memory (two barriers per
// round) and only chains 1024-element chunks of ONE block by look-back.
//
// Here a single kernel covers every (size, block_size, exclusive, reverse):
// a *segmented* single-pass scan.  Work is expressed in a "logical" index
// space I (I = i for forward scans, I = ntiles*TILE-1-i for reverse ones) in
// which block boundaries are the positions with (I + off) % block_size == 0.
// Because segments are regular, no head flags travel through the scan: every
// combine step is guarded by the distance to the most recent segment head,
// which each thread derives arithmetically.
//
//   - each thread owns J 16-byte vectors, loaded/stored fully coalesced
//     (vector v of the tile belongs to thread v % 256, slot v / 256);
//   - MODE 0: segments never cross a 32-vector warp row -> warp shuffles only;
//   - MODE 1: segments never cross a tile -> + one CTA-level combine;
//   - MODE 2: general: tiles are handed out by an atomic ticket (so every
//     predecessor of a waiting tile is already running) and chained with
//     decoupled look-back over 64-bit {status, value} descriptors.  A tile
//     that contains a segment head publishes its aggregate directly as an
//     inclusive prefix, which is also what stops look-back at block borders.
#include "scan.cuh"

namespace b200 {

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_WARPS = SCAN_THREADS / 32;
static constexpr uint32_t SCAN_CLAMP = 1u << 24; // > any tile size

struct ScanParams {
    const void *in;
    void *out;
    uint64_t size;      // elements
    uint32_t bs;        // segment length in logical space
    uint32_t bs_mask;   // bs - 1 if bs is a power of two, else 0
    uint32_t off;       // segment phase: s(I) = (I + off) % bs
    uint32_t ntiles;
    uint32_t exclusive;
    uint32_t reverse;
    uint64_t *desc;     // MODE 2: ntiles descriptors (zero-initialised)
    uint32_t *ticket;   // MODE 2: tile ticket counter (zero-initialised)
    const void *carry_in;
    void *carry_out;
    uint32_t vec_in;    // 'in' is 16-byte aligned: use 128-bit loads
    uint32_t vec_out;   // 'out' is 16-byte aligned: use 128-bit stores
};

template <typename T, int Op, int J, int MODE>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const ScanParams p) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr uint32_t TILE = SCAN_THREADS * J * N;
    constexpr int ENTRIES = J * SCAN_WARPS;
    static_assert(ENTRIES <= 32, "entry scan is done by one warp");

    __shared__ V s_incl[ENTRIES];
    __shared__ V s_excl[ENTRIES];
    __shared__ uint32_t s_t[ENTRIES];
    __shared__ uint32_t s_tile_id, s_tile_s;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T *in = (const T *) p.in;
    T *out = (T *) p.out;

    // ---- tile assignment and segment phase of the tile start
    uint32_t tile, s_tile;
    if constexpr (MODE == 2) {
        if (tid == 0) {
            uint32_t t = atomicAdd(p.ticket, 1u);
            s_tile_id = t;
            s_tile_s = (uint32_t) (((uint64_t) t * TILE + p.off) % p.bs);
        }
        __syncthreads();
        tile = s_tile_id;
        s_tile = s_tile_s;
    } else {
        tile = blockIdx.x;
        s_tile = 0; // TILE % bs == 0 and off == 0
    }

    // ---- load J vectors, local segmented inclusive scan
    V x[J][N];
    uint32_t s[J];     // #elements between the last head and the vector start
    uint32_t t[J];     // #elements from the last head through the vector end
    uint32_t heads[J]; // bit k: element k starts a segment
    uint64_t pbase[J]; // physical index of the vector's first (physical) element

    #pragma unroll
    for (int j = 0; j < J; ++j) {
        const uint32_t lv = j * SCAN_THREADS + tid; // logical vector in tile
        uint64_t sv = (uint64_t) s_tile + (uint64_t) lv * N;
        if (p.bs_mask)
            sv &= p.bs_mask;
        else if (p.bs >= TILE)
            sv = sv >= p.bs ? sv - p.bs : sv;
        else
            sv = (uint32_t) sv % p.bs;
        uint32_t pos = (uint32_t) sv;

        const uint32_t pv = p.reverse ? (SCAN_THREADS * J - 1) - lv : lv;
        const uint64_t ptile = p.reverse ? (p.ntiles - 1 - tile) : tile;
        const uint64_t pb = ptile * TILE + (uint64_t) pv * N;
        pbase[j] = pb;

        Vec16<T> v;
        if (p.vec_in && pb + N <= p.size) {
            v.raw = ld_stream_coherent(in + pb);
        } else {
            #pragma unroll
            for (int k = 0; k < N; ++k)
                if (pb + k < p.size)
                    v.elem[k] = in[pb + k];
        }

        uint32_t hm = 0, since = min(pos, SCAN_CLAMP);
        V run = R::identity();
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            const int pk = p.reverse ? N - 1 - k : k;
            V val = (pb + pk < p.size) ? to_value<T>(p.reverse ? v.elem[N - 1 - k] : v.elem[k])
                                       : R::identity();
            bool head = pos == 0;
            run = (k == 0 || head) ? val : R::apply(run, val);
            x[j][k] = run;
            hm |= head ? (1u << k) : 0u;
            since = head ? 1u : since + 1u;
            pos = (pos + 1 == p.bs) ? 0 : pos + 1;
        }
        s[j] = (uint32_t) min(sv, (uint64_t) SCAN_CLAMP);
        t[j] = since;
        heads[j] = hm;
    }

    // ---- warp-level segmented inclusive scan of the vector aggregates
    V a[J];
    #pragma unroll
    for (int j = 0; j < J; ++j)
        a[j] = x[j][N - 1];

    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            V up = shfl_up(a[j], d);
            if (lane >= d && t[j] > (uint32_t) (d * N))
                a[j] = R::apply(up, a[j]);
        }
    }

    V lane_excl[J];
    #pragma unroll
    for (int j = 0; j < J; ++j)
        lane_excl[j] = shfl_up(a[j], 1);

    // ---- CTA level: combine the J * WARPS row aggregates (one warp), chain tiles
    if constexpr (MODE >= 1) {
        if (lane == 31) {
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                s_incl[j * SCAN_WARPS + warp] = a[j];
                s_t[j * SCAN_WARPS + warp] = t[j];
            }
        }
        __syncthreads();

        if (warp == 0) {
            constexpr uint32_t ROW = 32 * N; // elements per entry
            const uint32_t m = lane;
            V A = m < ENTRIES ? s_incl[m] : R::identity();
            uint32_t tm = m < ENTRIES ? s_t[m] : SCAN_CLAMP;
            #pragma unroll
            for (int d = 1; d < ENTRIES; d <<= 1) {
                V up = shfl_up(A, d);
                if (m >= (uint32_t) d && tm > (uint32_t) d * ROW)
                    A = R::apply(up, A);
            }
            // segment phase of the entry's first element
            uint64_t sm64 = (uint64_t) s_tile + (uint64_t) m * ROW;
            if (p.bs_mask)
                sm64 &= p.bs_mask;
            else if (p.bs >= TILE)
                sm64 = sm64 >= p.bs ? sm64 - p.bs : sm64;
            else
                sm64 = (uint32_t) sm64 % p.bs;
            const uint32_t sm = (uint32_t) min(sm64, (uint64_t) SCAN_CLAMP);

            V prev = shfl_up(A, 1);
            V X = (m > 0 && sm > 0) ? prev : R::identity();

            if constexpr (MODE == 2) {
                const V tile_incl = shfl_idx(A, ENTRIES - 1);
                const uint32_t t_tile = __shfl_sync(FULL_MASK, tm, ENTRIES - 1);
                const bool has_head = t_tile <= TILE;
                const bool needs_prefix = s_tile > 0; // segment continues from before the tile
                V P = R::identity();

                if (!needs_prefix || has_head) {
                    // aggregate already is a complete inclusive prefix for successors
                    if (lane == 0)
                        Desc<V>::publish(p.desc, tile, DESC_PREFIX, tile_incl);
                } else if (tile > 0) {
                    if (lane == 0)
                        Desc<V>::publish(p.desc, tile, DESC_AGGREGATE, tile_incl);
                }

                if (needs_prefix) {
                    if (tile == 0) {
                        // only reachable with a carry-in (off == 1)
                        if (p.carry_in)
                            P = *(const V *) p.carry_in;
                    } else {
                        int64_t base = (int64_t) tile - 1;
                        while (true) {
                            int64_t idx = base - lane;
                            V val = R::identity();
                            uint32_t st = DESC_PREFIX;
                            do {
                                if (idx >= 0)
                                    st = Desc<V>::observe(p.desc, (uint32_t) idx, val);
                            } while (__any_sync(FULL_MASK, st == DESC_INVALID));
                            uint32_t ballot = __ballot_sync(FULL_MASK, st == DESC_PREFIX);
                            if (ballot) {
                                uint32_t first = __ffs(ballot) - 1;
                                V contrib = lane <= first ? val : R::identity();
                                P = R::apply(warp_reduce<V, Op>(contrib), P);
                                break;
                            }
                            P = R::apply(warp_reduce<V, Op>(val), P);
                            base -= 32;
                        }
                    }
                    if (!has_head && lane == 0)
                        Desc<V>::publish(p.desc, tile, DESC_PREFIX, R::apply(P, tile_incl));
                }

                if (sm > m * ROW)
                    X = R::apply(P, X);

                if (p.carry_out && tile == p.ntiles - 1 && lane == 0)
                    *(V *) p.carry_out = (needs_prefix && !has_head) ? R::apply(P, tile_incl)
                                                                     : tile_incl;
            }
            if (m < ENTRIES)
                s_excl[m] = X;
        }
        __syncthreads();
    }

    // ---- apply carries and store
    #pragma unroll
    for (int j = 0; j < J; ++j) {
        V carry = R::identity();
        if constexpr (MODE >= 1) {
            if (s[j] > lane * N)
                carry = s_excl[j * SCAN_WARPS + warp];
        }
        if (lane > 0 && s[j] > 0)
            carry = R::apply(carry, lane_excl[j]);

        const uint32_t hm = heads[j];
        const int first_head = hm ? __ffs(hm) - 1 : N;
        V incl[N], res[N];
        #pragma unroll
        for (int k = 0; k < N; ++k)
            incl[k] = k < first_head ? R::apply(carry, x[j][k]) : x[j][k];
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            if (p.exclusive) {
                V before = carry;
                if constexpr (true) {
                    if (k > 0)
                        before = incl[k > 0 ? k - 1 : 0];
                }
                res[k] = ((hm >> k) & 1u) ? R::identity() : before;
            } else {
                res[k] = incl[k];
            }
        }

        const uint64_t pb = pbase[j];
        if (p.vec_out && pb + N <= p.size) {
            Vec16<T> v;
            #pragma unroll
            for (int k = 0; k < N; ++k)
                v.elem[k] = from_value<T>(p.reverse ? res[N - 1 - k] : res[k]);
            st_stream(out + pb, v.raw);
        } else {
            #pragma unroll
            for (int k = 0; k < N; ++k)
                if (pb + k < p.size)
                    out[pb + k] = from_value<T>(p.reverse ? res[N - 1 - k] : res[k]);
        }
    }
}

// ---------------------------------------------------------------- dispatch

template <typename T, int Op> static int launch_scan(const ScanCall &c) {
    using V = typename ValueOf<T>::type;
    constexpr int N = VecInfo<T>::N;
    constexpr int J = 4;
    constexpr uint32_t TILE = SCAN_THREADS * J * N;

    uint64_t ntiles64 = ceil_div(c.size, TILE);
    if (c.size >= 0xffffffffull - 2 * TILE)
        return fail(B200_ERR_INVALID, "jit_block_prefix_reduce(): array too large!");
    uint32_t ntiles = (uint32_t) ntiles64;

    ScanParams p{};
    p.in = c.in;
    p.out = c.out;
    p.size = c.size;
    p.ntiles = ntiles;
    p.exclusive = c.exclusive;
    p.reverse = c.reverse;
    p.carry_in = c.carry_in;
    p.carry_out = c.carry_out;
    // misaligned arrays fall back to element-wise (still coalesced) accesses
    p.vec_in = ((uintptr_t) c.in % 16) == 0;
    p.vec_out = ((uintptr_t) c.out % 16) == 0;

    int mode;
    if (c.carry_api) {
        p.bs = 0xffffffffu;
        p.off = c.carry_in ? 1 : 0;
        mode = 2;
    } else {
        p.bs = (uint32_t) c.bs;
        p.off = c.reverse ? (uint32_t) ((c.bs - (ntiles64 * TILE) % c.bs) % c.bs) : 0;
        if ((32 * N) % c.bs == 0)
            mode = 0;
        else if (TILE % c.bs == 0)
            mode = 1;
        else
            mode = 2;
    }
    p.bs_mask = is_pow2(p.bs) ? p.bs - 1 : 0;

    void *scratch = nullptr;
    if (mode == 2) {
        size_t desc_bytes = (size_t) ntiles * Desc<V>::WORDS * sizeof(uint64_t);
        scratch = temp_alloc(desc_bytes + 16, c.stream);
        if (!scratch)
            return fail(B200_ERR_CUDA, "jit_block_prefix_reduce(): out of memory");
        B200_CUDA_CHECK(cudaMemsetAsync(scratch, 0, desc_bytes + 16, c.stream));
        p.desc = (uint64_t *) scratch;
        p.ticket = (uint32_t *) ((uint8_t *) scratch + desc_bytes);
    }

    switch (mode) {
        case 0: scan_kernel<T, Op, J, 0><<<ntiles, SCAN_THREADS, 0, c.stream>>>(p); break;
        case 1: scan_kernel<T, Op, J, 1><<<ntiles, SCAN_THREADS, 0, c.stream>>>(p); break;
        default: scan_kernel<T, Op, J, 2><<<ntiles, SCAN_THREADS, 0, c.stream>>>(p); break;
    }
    if (scratch)
        temp_free(scratch, c.stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

typedef int (*ScanFn)(const ScanCall &);

template <typename T, bool Bits> static ScanFn pick_scan_op(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_scan<T, B200_OP_ADD>;
        case B200_OP_MUL: return launch_scan<T, B200_OP_MUL>;
        case B200_OP_MIN: return launch_scan<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_scan<T, B200_OP_MAX>;
        case B200_OP_AND: if constexpr (Bits) return launch_scan<T, B200_OP_AND>; else return nullptr;
        case B200_OP_OR:  if constexpr (Bits) return launch_scan<T, B200_OP_OR>; else return nullptr;
        default: return nullptr;
    }
}

template <typename T> static ScanFn pick_scan_minmax(int op) {
    switch (op) {
        case B200_OP_MIN: return launch_scan<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_scan<T, B200_OP_MAX>;
        default: return nullptr;
    }
}

static ScanFn pick_scan(int vt, int op) {
    bool minmax = op == B200_OP_MIN || op == B200_OP_MAX;
    switch (vt) {
        case B200_VT_INT32:  return minmax ? pick_scan_minmax<int32_t>(op) : pick_scan_op<uint32_t, true>(op);
        case B200_VT_UINT32: return pick_scan_op<uint32_t, true>(op);
        case B200_VT_INT64:  return minmax ? pick_scan_minmax<int64_t>(op) : pick_scan_op<uint64_t, true>(op);
        case B200_VT_UINT64: return pick_scan_op<uint64_t, true>(op);
        case B200_VT_FLOAT16: return pick_scan_op<__half, false>(op);
        case B200_VT_FLOAT32: return pick_scan_op<float, false>(op);
        case B200_VT_FLOAT64: return pick_scan_op<double, false>(op);
        default: return nullptr;
    }
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_block_prefix_reduce(void *stream_, int vt, int op, uint64_t size,
                             uint64_t block_size, int exclusive, int reverse,
                             const void *in, void *out) {
    int rc = ensure_init();
    if (rc)
        return rc;
    // src/cuda_ts.cpp:539-552
    if (size == 0)
        return B200_OK;
    if (block_size == 0 || block_size > size)
        return fail(B200_ERR_INVALID,
                    "jit_block_prefix_reduce(): invalid block size (size=%llu, block_size=%llu)!",
                    (unsigned long long) size, (unsigned long long) block_size);
    uint32_t tsize = type_size(vt);
    ScanFn fn = pick_scan(vt, op);
    if (!fn || tsize == 0)
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
                    type_name(vt), op_name(op));
    if (((uintptr_t) in % tsize) != 0 || ((uintptr_t) out % tsize) != 0)
        return fail(B200_ERR_INVALID, "jit_block_prefix_reduce(): misaligned pointer!");
    cudaStream_t stream = resolve_stream(stream_);
    HistoryScope hs(stream, B200_KERNEL_BLOCK_PREFIX_REDUCE, size);
    if (block_size == 1) {
        if (exclusive) {
            uint64_t ident = b200_reduce_identity(vt, op);
            return b200_memset_async(stream, out, size, tsize, &ident);
        } else if (in != out) {
            B200_CUDA_CHECK(cudaMemcpyAsync(out, in, size * tsize, cudaMemcpyDeviceToDevice, stream));
        }
        return B200_OK;
    }
    ScanCall call{ stream, in, out, size, block_size, exclusive != 0, reverse != 0,
                   nullptr, nullptr, false };
    bool handled = false;
    rc = scan_fast_dispatch(vt, op, call, &handled);
    if (rc || handled)
        return rc;
    return fn(call);
}

int b200_prefix_reduce_carry(void *stream_, int vt, int op, uint64_t size, int exclusive,
                             int reverse, const void *in, void *out, const void *carry_in,
                             void *carry_out) {
    int rc = ensure_init();
    if (rc)
        return rc;
    uint32_t tsize = type_size(vt);
    ScanFn fn = pick_scan(vt, op);
    if (!fn || tsize == 0)
        return fail(B200_ERR_UNSUPPORTED,
                    "jit_block_prefix_reduce(): no existing kernel for type=%s, op=%s!",
                    type_name(vt), op_name(op));
    cudaStream_t stream = resolve_stream(stream_);
    if (size == 0) {
        if (carry_out) {
            uint32_t vsize = vt == B200_VT_FLOAT16 ? 4 : tsize;
            if (carry_in) {
                B200_CUDA_CHECK(cudaMemcpyAsync(carry_out, carry_in, vsize, cudaMemcpyDeviceToDevice, stream));
            } else {
                return fail(B200_ERR_INVALID, "b200_prefix_reduce_carry(): empty array without carry!");
            }
        }
        return B200_OK;
    }
    ScanCall call{ stream, in, out, size, size, exclusive != 0, reverse != 0, carry_in,
                   carry_out, true };
    bool handled = false;
    rc = scan_fast_dispatch(vt, op, call, &handled);
    if (rc || handled)
        return rc;
    return fn(call);
}

uint32_t b200_scan_tile_elems(int vt) {
    const uint32_t tsize = type_size(vt);
    return tsize ? 32768u / tsize : 0u; // DefaultGeom of scan_fast.cu: 512 threads x 4 x 16 bytes
}

int b200_prefix_reduce_seeded(void *stream_, int vt, int op, uint64_t size, int exclusive,
                              int reverse, const void *in, void *out, const void *tile_seeds) {
    int rc = ensure_init();
    if (rc)
        return rc;
    const uint32_t tsize = type_size(vt);
    if (!pick_scan(vt, op) || tsize == 0 || vt == B200_VT_FLOAT16)
        return fail(B200_ERR_UNSUPPORTED,
                    "b200_prefix_reduce_seeded(): no existing kernel for type=%s, op=%s!",
                    type_name(vt), op_name(op));
    if (size == 0)
        return B200_OK;
    if (!tile_seeds)
        return fail(B200_ERR_INVALID, "b200_prefix_reduce_seeded(): tile_seeds is NULL!");
    ScanCall call{ resolve_stream(stream_), in, out, size, size, exclusive != 0, reverse != 0,
                   nullptr, nullptr, true };
    call.seeds = tile_seeds;
    bool handled = false;
    rc = scan_fast_dispatch(vt, op, call, &handled);
    if (rc)
        return rc;
    if (!handled)
        return fail(B200_ERR_UNSUPPORTED,
                    "b200_prefix_reduce_seeded(): needs 16-byte aligned arrays of fewer than 2^32 - 2^14 "
                    "elements (use b200_prefix_reduce_carry otherwise)");
    return B200_OK;
}

} // extern "C"
