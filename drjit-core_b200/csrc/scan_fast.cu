// scan_fast.cu -- streaming prefix reductions: the fast paths of
// jit_block_prefix_reduce (include/drjit-core/jit.h:2365-2373).
//
// Replaces the kernel of resources/block_prefix_reduce.cuh (one element per
// thread, Hillis-Steele in shared memory with two barriers per round, look-back
// only between the 1024-element chunks of one block) for the two shapes that
// matter at scale:
//
//   POW2   block_size is a power of two and at most one tile: blocks never
//          cross a tile, so tiles are independent.
//   CHAIN  the whole array is one block (optionally seeded with a carry), or
//          block_size is a power-of-two multiple of the tile: tiles are chained
//          with decoupled look-back over 64-bit {status, value} descriptors.
//          Blocks of 2..16 tiles are scanned by one CTA each (tile groups: no global
//          look-back); whole arrays and larger blocks of 4-byte types, from 1024 tiles
//          on, run as two streams (scan_ahead_kernel below: a reduce stream ahead of
//          the scan stream, one prefix stream over the tile aggregates).
//   SEG    any other block size: CHAIN with block boundaries at arbitrary elements.
//          Boundaries are regular, so no head flags travel through the scan: every
//          combine step is guarded by the distance of a vector to the most recent
//          block start, which each thread derives arithmetically.  A tile that
//          contains a block start publishes the aggregate of its tail as an inclusive
//          prefix at once, which is also what stops the look-back at block borders.
//   In all modes tiles are handed out by an atomic ticket: every predecessor
//   of a tile is therefore owned by a CTA that is already running (forward
//   progress of the look-back), and fast SMs take more tiles (load balance).
//
// Pointers that are not 16-byte aligned (and the carry / seeded entry points with
// such pointers) are served by the general segmented kernel in scan.cu.
//
// Kernel structure (persistent CTAs, 512 compute threads):
//   - a tile is 512 * J 16-byte vectors (32 KiB for J = 4).  Thread 0 keeps a
//     ring of S shared-memory slots filled with 1-D bulk copies (the TMA engine,
//     cp.async.bulk + mbarrier complete_tx) S - 2 tiles ahead of the tile being
//     computed, so loads stay in flight while the CTA computes;
//   - CHAIN only: two helper warps run ahead of the compute warps on the tiles
//     that have landed in the ring.  One reduces the tile and publishes its
//     aggregate immediately, the other resolves the tile's exclusive prefix by
//     look-back and hands it over through an mbarrier, so neither the look-back
//     latency nor a neighbour waiting for this CTA's aggregate stalls the
//     stream;
//   - warp w owns the contiguous rows [w * J, (w + 1) * J) of 32 vectors; a
//     thread scans its vector in registers, rows are scanned with shuffles
//     (only log2(lanes per block) steps for small blocks), row and warp totals
//     are combined through registers / one shared-memory exchange;
//   - results are written back into the slot and leave with one bulk store per
//     tile (cp.async.bulk.global.shared::cta), which also drains asynchronously.
//   The last, partial tile of an array uses guarded 128-bit loads / stores.
#include "scan.cuh"
#include "pipeline.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>

#ifndef B200_SCAN_LBW
#define B200_SCAN_LBW 1
#endif

namespace b200 {

template <int X> struct CLog2 { static constexpr int value = X <= 1 ? 0 : 1 + CLog2<X / 2>::value; };

struct FastParams {
    const void *in;
    void *out;
    uint64_t size;       // elements
    uint32_t ntiles;
    uint32_t log2_bs;    // POW2: log2(block_size)
    uint32_t seg_mask;   // CHAIN: tiles per block - 1 (0xffffffff: one block)
    uint32_t tile_off;   // CHAIN: phase of the block grid in logical tile space
    uint32_t lag;        // AHEAD: tiles the reduce stream may run in front of the scan stream
    uint32_t log2_group; // CHAIN: a ticket stands for 2^log2_group neighbouring tiles (one global look-back per group)
    uint32_t bs;         // SEG: block size; logical element I starts a block iff (I + off) % bs == 0
    uint32_t off;        // SEG: phase of the block grid in logical element space
    uint32_t exclusive;
    uint32_t reverse;
    uint32_t debug;      // development switches
    uint64_t *desc;      // CHAIN: ntiles descriptors (zero-initialised)
    uint32_t *ticket;    // tile ticket counter (zero-initialised)
    uint32_t *done;      // optional: CTAs that have finished; the last one zeroes both counters
    const void *carry_in;
    void *carry_out;
    const void *seeds;   // CHAIN: exclusive prefix per PHYSICAL tile, replaces the look-back
    CyclicScan cyc;      // CHAIN: block-cyclic multi-GPU scan (world == 0: off)
};

/// 16-byte {value, epoch} entries exchanged between GPUs: one single-copy atomic
/// 128-bit access at system scope, so a reader that sees the epoch sees its value
B200_DEVICE void st_entry_sys(uint64_t *ptr, uint64_t value, uint64_t epoch) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.sys.global.b128 [%0], q;\n\t}"
                 :: "l"(ptr), "l"(value), "l"(epoch) : "memory");
}
B200_DEVICE void ld_entry_sys(const uint64_t *ptr, uint64_t &value, uint64_t &epoch) {
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.sys.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(value), "=l"(epoch) : "l"(ptr) : "memory");
}
/// Wait (bounded: 20 s) until the entry carries `epoch`; false on time-out
B200_DEVICE bool wait_entry_sys(const uint64_t *ptr, uint64_t epoch, uint64_t &value) {
    uint64_t e, t0 = 0;
    uint32_t spins = 0;
    while (true) {
        ld_entry_sys(ptr, value, e);
        if (e == epoch)
            return true;
        if ((++spins & 255u) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > 20ull * 1000 * 1000 * 1000)
                return false;
        }
        __nanosleep(64);
    }
}

#if defined(B200_SCAN_TUNING)
// development counters (accumulated in registers, flushed once per warp):
// [0] look-back steps, [1] polls, [2] cycles in look-back, [3] compute cycles
// waiting for a prefix, [4] compute cycles waiting for a tile to land, [5] tiles,
// [6] aggregate-warp cycles waiting for a tile, [7] look-back cycles waiting for
// the aggregate, [8] aggregate-warp cycles reducing, [9] thread-0 cycles in
// issue(), [10] compute-loop cycles in total (thread 32), [11] thread-0 cycles
// waiting for the aggregate warp before a refill
__device__ unsigned long long g_scan_dbg[16];
#define DBG_DECL() unsigned long long dbg_acc[16] = { 0 }
#define DBG_ADD(i, v) dbg_acc[i] += (unsigned long long) (v)
#define DBG_ADD_LANE0(i, v) DBG_ADD(i, v)
#define DBG_FLUSH() do { if (p.debug & 2) { for (int i_ = 0; i_ < 16; ++i_) if (dbg_acc[i_]) atomicAdd(&g_scan_dbg[i_], dbg_acc[i_]); } } while (0)
#define DBG_CLOCK() clock64()
#else
#define DBG_DECL() ((void) 0)
#define DBG_ADD(i, v) ((void) 0)
#define DBG_ADD_LANE0(i, v) ((void) 0)
#define DBG_FLUSH() ((void) 0)
#define DBG_CLOCK() 0ll
#endif

/// Barrier among the THREADS compute threads only (the look-back warp of the
/// CHAIN kernel never joins it)
template <int THREADS> B200_DEVICE void compute_sync() {
    asm volatile("bar.sync 1, %0;" :: "n"(THREADS) : "memory");
}

/// Scan of one tile held in registers.  On entry e[j][k] holds the tile's values
/// in logical order (warp-contiguous rows); on exit e[j][k] is the inclusive
/// prefix inside the vector and carry[j] the exclusive prefix of the vector
/// within its block *inside this tile*.  FULL: the tile is one run (no block
/// boundaries).
template <typename T, int Op, int J, int THREADS, bool FULL> struct TileScan {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    static constexpr int N = VecInfo<T>::N;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int LOG2_N = CLog2<N>::value;
    static constexpr int LOG2_J = CLog2<J>::value;
    static constexpr int LOG2_W = CLog2<WARPS>::value;

    static B200_DEVICE void run(V (&e)[J][N], V (&carry)[J], uint32_t log2_bs, uint32_t lane,
                                uint32_t warp, V *s_warp) {
        // ---- inside the vector
        const uint32_t mask = FULL ? 0xffffffffu : ((1u << log2_bs) - 1u);
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            #pragma unroll
            for (int k = 1; k < N; ++k)
                if (FULL || (k & mask))
                    e[j][k] = R::apply(e[j][k - 1], e[j][k]);
        }

        // ---- inside a row: blocks of L lanes
        uint32_t L = 32;
        if (!FULL) {
            L = (1u << log2_bs) >> LOG2_N;
            L = L < 1 ? 1 : (L > 32 ? 32 : L);
        }
        const uint32_t sl = lane & (L - 1);
        V a[J];
        #pragma unroll
        for (int j = 0; j < J; ++j)
            a[j] = e[j][N - 1];
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (FULL || (uint32_t) d < L) {
                #pragma unroll
                for (int j = 0; j < J; ++j) {
                    V up = shfl_up(a[j], d);
                    if (sl >= (uint32_t) d)
                        a[j] = R::apply(up, a[j]);
                }
            }
        }
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            V up = shfl_up(a[j], 1);
            carry[j] = sl ? up : R::identity();
        }

        // ---- across the rows of a warp, then across warps
        if (FULL || log2_bs > (uint32_t) (LOG2_N + 5)) {
            uint32_t rows = J;
            if (!FULL) {
                uint32_t lr = log2_bs - (LOG2_N + 5);
                rows = lr >= (uint32_t) LOG2_J ? J : (1u << lr);
            }
            V run = R::identity();
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                V rt = shfl_idx(a[j], 31);
                if (!FULL && (j & (rows - 1)) == 0)
                    run = R::identity();
                carry[j] = R::apply(run, carry[j]);
                run = R::apply(run, rt);
            }

            if (FULL || log2_bs > (uint32_t) (LOG2_N + 5 + LOG2_J)) {
                uint32_t warps = WARPS;
                if (!FULL) {
                    uint32_t lw = log2_bs - (LOG2_N + 5 + LOG2_J);
                    warps = lw >= (uint32_t) LOG2_W ? WARPS : (1u << lw);
                }
                if (lane == 0)
                    s_warp[warp] = run;
                compute_sync<THREADS>();
                const uint32_t first = warp & ~(warps - 1);
                V wc = R::identity();
                #pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    V t = s_warp[w];
                    if ((uint32_t) w >= first && (uint32_t) w < warp)
                        wc = R::apply(wc, t);
                }
                #pragma unroll
                for (int j = 0; j < J; ++j)
                    carry[j] = R::apply(wc, carry[j]);
            }
        }
    }
};

/// Position inside its block of the element `delta` (< bs) elements after one at position `pos` (< bs)
B200_DEVICE uint32_t seg_advance(uint32_t pos, uint32_t delta, uint32_t bs) {
    return delta >= bs - pos ? delta - (bs - pos) : pos + delta;
}

/// TileScan for blocks of ANY size `bs` >= 2 (SEG).  sv[j]: position of the vector's first
/// element inside its block (0: it starts a block).  On exit e[j][k] is the inclusive prefix
/// since the nearest block start INSIDE the vector (or since the vector's start), carry[j] the
/// reduction of the elements of this tile in front of the vector that belong to the block of
/// its first element.  It applies to the elements in front of the vector's first block start.
/// ONE: bs >= N, i.e. a vector contains at most one block start -- element h[j] (N: none).
template <typename T, int Op, int J, int THREADS, bool ONE> struct TileScanSeg {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    static constexpr int N = VecInfo<T>::N;
    static constexpr int WARPS = THREADS / 32;

    /// First block start inside a vector whose first element sits at position `sv` (ONE)
    static B200_DEVICE uint32_t first_head(uint32_t sv, uint32_t bs) {
        return sv == 0 ? 0u : min(bs - sv, (uint32_t) N);
    }

    static B200_DEVICE void run(V (&e)[J][N], V (&carry)[J], const uint32_t (&sv)[J], uint32_t bs, uint32_t lane,
                                uint32_t warp, V *s_warp, uint32_t *s_wt) {
        // ---- inside the vector; t[j]: elements from the last block start through the vector's end
        uint32_t t[J];
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            if constexpr (ONE) {
                const uint32_t h = first_head(sv[j], bs);
                #pragma unroll
                for (int k = 1; k < N; ++k)
                    if ((uint32_t) k != h)
                        e[j][k] = R::apply(e[j][k - 1], e[j][k]);
                t[j] = h < (uint32_t) N ? N - h : sv[j] + N;
            } else {
                uint32_t pos = sv[j];
                #pragma unroll
                for (int k = 1; k < N; ++k) {
                    pos = pos + 1 == bs ? 0u : pos + 1;
                    if (pos != 0)
                        e[j][k] = R::apply(e[j][k - 1], e[j][k]);
                }
                t[j] = pos + 1;
            }
        }
        // ---- inside a row: the vector d lanes down belongs to my block iff t > d * N,
        // i.e. iff d <= reach = min(lane, (t - 1) / N)
        V a[J];
        uint32_t reach[J];
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            a[j] = e[j][N - 1];
            reach[j] = min(lane, (t[j] - 1) / (uint32_t) N);
        }
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                V up = shfl_up(a[j], d);
                if (reach[j] >= (uint32_t) d)
                    a[j] = R::apply(up, a[j]);
            }
        }
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            V up = shfl_up(a[j], 1);
            carry[j] = lane && sv[j] ? up : R::identity();
        }
        // ---- across the rows of a warp: `run` = the tail of the rows so far
        V run = R::identity();
        uint32_t tr = 0;
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const V rt = shfl_idx(a[j], 31);
            tr = __shfl_sync(FULL_MASK, t[j], 31);
            if (sv[j] > lane * N) // the vector's block starts in front of this row
                carry[j] = R::apply(run, carry[j]);
            run = tr > 32u * N ? R::apply(run, rt) : rt;
        }
        // ---- across warps: the tail of the last warp in front of this one that contains a
        // block start, and all of the warps between the two
        if (lane == 0) {
            s_warp[warp] = run;
            s_wt[warp] = tr;
        }
        compute_sync<THREADS>();
        static_assert(WARPS <= 32, "one lane per warp");
        const bool mine = lane < warp;
        const V wv = mine ? s_warp[lane] : R::identity();
        const uint32_t heads = __ballot_sync(FULL_MASK, mine && s_wt[lane] <= (uint32_t) (J * 32 * N));
        const uint32_t from = heads ? 31u - (uint32_t) __clz((int) heads) : 0u;
        const V wc = warp_reduce<V, Op>(lane >= from ? wv : R::identity());
        #pragma unroll
        for (int j = 0; j < J; ++j)
            if (sv[j] > (j * 32 + lane) * N) // ... in front of this warp's rows
                carry[j] = R::apply(wc, carry[j]);
    }

    /// Results of vector j from e / carry (after the tile's prefix has been folded into carry)
    static B200_DEVICE void results(const V (&e)[N], V carry, uint32_t sv, uint32_t bs, bool exclusive, V (&res)[N]) {
        // `pre`: what precedes the vector within the block of its first element; it ends at the
        // first block start inside the vector
        if constexpr (ONE) {
            // elements in front of the block start continue `carry`; an exclusive result is the
            // inclusive one of the element before, or the identity at the block start
            const uint32_t h = first_head(sv, bs);
            V incl[N];
            #pragma unroll
            for (int kk = 0; kk < N; ++kk)
                incl[kk] = (uint32_t) kk < h ? R::apply(carry, e[kk]) : e[kk];
            #pragma unroll
            for (int kk = 0; kk < N; ++kk) {
                if (!exclusive)
                    res[kk] = incl[kk];
                else
                    res[kk] = (uint32_t) kk == h ? R::identity() : (kk == 0 ? carry : incl[kk - 1]);
            }
        } else {
            V pre = sv ? carry : R::identity();
            uint32_t pos = sv;
            #pragma unroll
            for (int kk = 0; kk < N; ++kk) {
                if (pos == 0)
                    pre = R::identity();
                if (exclusive)
                    res[kk] = pos == 0 ? R::identity() : (kk == 0 ? pre : R::apply(pre, e[kk - 1]));
                else
                    res[kk] = R::apply(pre, e[kk]);
                pos = pos + 1 == bs ? 0u : pos + 1;
            }
        }
    }
};

/// THREADS compute threads (+ one look-back warp when CHAIN), tiles of
/// THREADS * J vectors, ring of S shared-memory slots.
template <typename T, int Op, int J, int S, int THREADS, bool CHAIN, int LB, int LBW, bool SEG = false>
__global__ void __launch_bounds__(THREADS + (CHAIN ? 32 + 32 * LB : 0),
                                  (THREADS >= 512 && (size_t) S * THREADS * J * 16 <= 112 * 1024) ? 2 : 1)
scan_stream_kernel(const FastParams p) {
    static_assert(CHAIN || !SEG, "SEG is a CHAIN mode");
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t VECS = THREADS * J;
    constexpr uint32_t TILE = VECS * N;
    constexpr uint32_t TILE_BYTES = VECS * 16;
    constexpr uint32_t NONE = 0xffffffffu;
    static_assert(LB <= S, "every look-back warp must meet one of the S trailing empty slots");

    extern __shared__ __align__(128) uint8_t sf_smem[];
    uint4 *slots = (uint4 *) sf_smem; // S * VECS vectors
    __shared__ __align__(8) uint64_t s_full[S];  // bulk load of the slot has landed
    // CHAIN: hand-over between aggregate warp -> look-back warps -> compute warps.
    // These rings have 2 * S entries (tile k uses entry k % M): a slot is handed
    // back to the producer as soon as the compute warps hold the tile in
    // registers, i.e. possibly before the look-back of that tile has started.
    constexpr int M = 2 * S;
    __shared__ __align__(8) uint64_t s_agg[M];   // s_total / s_mtile entry is valid
    __shared__ __align__(8) uint64_t s_pref[M];  // s_prefix entry is valid
    __shared__ uint32_t s_tile[S];
    __shared__ uint32_t s_spos[S];   // SEG: position of the tile's first logical element inside its block
    __shared__ uint32_t s_mtile[M];
    __shared__ uint32_t s_mseg[M];   // SEG: bit 0 = the tile contains a block start, bit 1 = it begins with one
    __shared__ uint32_t s_wt[WARPS];
    __shared__ V s_total[M];
    __shared__ V s_local[M];   // aggregate of the tiles of the same group in front of the tile
    __shared__ V s_prefix[M];
    __shared__ V s_warp[WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T *in = (const T *) p.in;
    T *out = (T *) p.out;
    DBG_DECL();

    // ---- producer side (thread 0): tile assignment + bulk loads into the ring
    // Tiles are handed out by an atomic ticket in both modes.  CHAIN needs it
    // for forward progress; POW2 for load balance (measured: with a static
    // round-robin assignment the slowest SM is busy 23 % longer than the
    // fastest one and the kernel loses 13 %).
    auto issue = [&](uint32_t slot, uint32_t t) {
        if (t >= p.ntiles) {
            s_tile[slot] = NONE;
            mbar_arrive(&s_full[slot]);
            return;
        }
        s_tile[slot] = t;
        if constexpr (SEG)
            s_spos[slot] = (uint32_t) (((uint64_t) t * TILE + p.off) % p.bs);
        const uint32_t pt = p.reverse ? p.ntiles - 1 - t : t;
        const uint64_t base = (uint64_t) pt * TILE;
        if (base + TILE <= p.size) {
            mbar_arrive_expect_tx(&s_full[slot], TILE_BYTES);
            bulk_g2s(slots + (size_t) slot * VECS, in + base, TILE_BYTES, &s_full[slot]);
        } else {
            mbar_arrive(&s_full[slot]); // partial tile: loaded directly by its readers
        }
    };

    // CHAIN: a ticket stands for a GROUP of 2^log2_group neighbouring tiles, which this CTA takes one after
    // the other.  Every tile still publishes its own aggregate the moment it has landed, but only the
    // group's first tile looks back through global memory: the others continue from the aggregates the
    // CTA already holds, so the look-back warp -- one global look-back per tile kept it busy 90 % of the
    // time -- stops being the bottleneck.  (One DESCRIPTOR per group was measured as well and is slower,
    // 0.47 vs 0.39 ms at 2^28: the group's aggregate is complete one ring turn later and every successor
    // polls that long.)
    const uint32_t glog = CHAIN ? p.log2_group : 0u, gmask = (1u << glog) - 1u;
    uint32_t tk_next = 0, tk_left = 0; // (thread 0)
    auto next_tile = [&]() -> uint32_t {
        if (tk_left == 0) {
            tk_next = atomicAdd(p.ticket, 1u) << glog;
            tk_left = gmask + 1;
        }
        --tk_left;
        return tk_next++;
    };

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < S; ++s)
            mbar_init(&s_full[s], 1);
        #pragma unroll
        for (int s = 0; s < M; ++s) {
            mbar_init(&s_agg[s], 1);
            mbar_init(&s_pref[s], 1);
        }
        mbar_fence_init();
        for (uint32_t i = 0; i < (uint32_t) S; ++i)
            issue(i, next_tile());
    }
    __syncthreads();

    // Exclusive prefix of `tile` within its segment by decoupled look-back (one warp)
    auto look_back = [&](uint32_t tile) -> V {
        V P = R::identity();
        int64_t win = (int64_t) tile - 1;
        bool done = false;
        while (!done) {
            // LBW windows of 32 preceding tiles each are fetched TOGETHER (their
            // loads are independent: one L2 round trip, ~0.9 us under streaming
            // load, covers 32 * LBW descriptors); lane 0 of window 0 is the
            // nearest tile.  The windows are then examined nearest first.
            // Polling competes with the stream for L2 bandwidth, so only lanes
            // whose entry is still INVALID poll again, after a short back-off;
            // entries beyond the nearest PREFIX are not needed at all.
            V val[LBW];
            uint32_t st[LBW];
            #pragma unroll
            for (int w = 0; w < LBW; ++w) {
                const int64_t idx = win - 32 * w - lane;
                val[w] = R::identity();
                st[w] = DESC_PREFIX;
                if (idx >= 0)
                    st[w] = Desc<V>::observe(p.desc, (uint32_t) idx, val[w]);
            }
            DBG_ADD_LANE0(0, 1);
            DBG_ADD_LANE0(1, 1);
            #pragma unroll
            for (int w = 0; w < LBW; ++w) {
                if (done)
                    break;
                const int64_t idx = win - 32 * w - lane;
                uint32_t pre;
                while (true) {
                    pre = __ballot_sync(FULL_MASK, st[w] == DESC_PREFIX);
                    uint32_t inv = __ballot_sync(FULL_MASK, st[w] == DESC_INVALID);
                    if (pre)
                        inv &= (1u << (__ffs(pre) - 1)) - 1u;
                    if (!inv)
                        break;
                    __nanosleep(100);
                    if (st[w] == DESC_INVALID)
                        st[w] = Desc<V>::observe(p.desc, (uint32_t) idx, val[w]);
                    DBG_ADD_LANE0(1, 1);
                }
                if (pre) {
                    const uint32_t stop = __ffs(pre) - 1;
                    V contrib = lane <= stop ? val[w] : R::identity();
                    P = R::apply(warp_reduce<V, Op>(contrib), P);
                    done = true;
                } else {
                    P = R::apply(warp_reduce<V, Op>(val[w]), P);
                }
            }
            win -= 32 * LBW;
        }
        return P;
    };

    if constexpr (CHAIN) {
        // ---- two helper warps run ahead of the compute warps.
        // Aggregate warp: reduces every tile as soon as it has landed and
        // publishes the aggregate at once -- never held up by a look-back, so
        // that the successors of a tile find its aggregate without waiting for
        // this CTA.
        if (warp == WARPS) {
            int none_seen = 0;
            V gacc = R::identity(); // lane 0: aggregate of the current group so far
            for (uint32_t k = 0;; ++k) {
                const uint32_t slot = k % S;
                long long c0 = DBG_CLOCK();
                mbar_wait(&s_full[slot], (k / S) & 1);
                long long c1 = DBG_CLOCK();
                DBG_ADD(6, c1 - c0);
                const uint32_t tile = s_tile[slot];
                const uint32_t m = k % M;
                if (tile == NONE) {
                    // wake every look-back warp: the next LB slots are NONE too
                    if (lane == 0) {
                        s_mtile[m] = NONE;
                        mbar_arrive(&s_agg[m]);
                    }
                    if (++none_seen == LB)
                        break;
                    continue;
                }
                (void) c1;
                if (p.seeds) { // prefixes are supplied: nothing to reduce or publish
                    if (lane == 0) {
                        s_mtile[m] = tile;
                        mbar_arrive(&s_agg[m]);
                    }
                    __syncwarp();
                    continue;
                }
                const uint32_t ptile = p.reverse ? p.ntiles - 1 - tile : tile;
                const uint64_t base = (uint64_t) ptile * TILE;
                const uint4 *slot_ptr = slots + (size_t) slot * VECS;
                constexpr int AU = 8; // vectors in flight per lane
                static_assert(VECS % (32 * AU) == 0, "tile must be a multiple of 256 vectors");
                V acc[AU];
                #pragma unroll
                for (int u = 0; u < AU; ++u)
                    acc[u] = R::identity();
                // SEG: only the tile's tail -- the elements from its last block start on --
                // reaches the tiles behind it: PHYSICAL elements [lo, hi) of the tile
                uint32_t lo = 0, hi = TILE, seg_bits = 0;
                if constexpr (SEG) {
                    const uint32_t spos = s_spos[slot];
                    const uint32_t t_end = seg_advance(spos, (TILE - 1) % p.bs, p.bs) + 1;
                    const bool has_head = p.bs <= TILE || t_end <= TILE;
                    seg_bits = (has_head ? 1u : 0u) | (spos == 0 ? 2u : 0u);
                    if (has_head) {
                        if (p.reverse)
                            hi = t_end;
                        else
                            lo = TILE - t_end;
                    }
                }
                if (base + TILE <= p.size) {
                    #pragma unroll 1
                    for (uint32_t i = SEG ? (lo / N / (32 * AU)) * (32 * AU) + lane : lane; i < VECS; i += 32 * AU) {
                        if (SEG && i - lane >= hi / N + 1)
                            break; // (warp-uniform)
                        Vec16<T> v[AU];
                        #pragma unroll
                        for (int u = 0; u < AU; ++u)
                            v[u].raw = slot_ptr[i + u * 32];
                        #pragma unroll
                        for (int u = 0; u < AU; ++u) {
                            // pairwise inside the vector: short dependency chains
                            V t[N];
                            #pragma unroll
                            for (int kk = 0; kk < N; ++kk)
                                t[kk] = to_value<T>(v[u].elem[kk]);
                            if constexpr (SEG) {
                                const uint32_t pe = (i + u * 32) * N;
                                if (pe < lo || pe + N > hi) {
                                    #pragma unroll
                                    for (int kk = 0; kk < N; ++kk)
                                        if (pe + kk < lo || pe + kk >= hi)
                                            t[kk] = R::identity();
                                }
                            }
                            #pragma unroll
                            for (int w = N / 2; w > 0; w >>= 1) {
                                #pragma unroll
                                for (int kk = 0; kk < w; ++kk)
                                    t[kk] = R::apply(t[kk], t[kk + w]);
                            }
                            acc[u] = R::apply(acc[u], t[0]);
                        }
                    }
                } else {
                    for (uint64_t i = base + lo + lane; i < p.size && i < base + hi; i += 32)
                        acc[0] = R::apply(acc[0], to_value<T>(in[i]));
                }
                #pragma unroll
                for (int w = AU / 2; w > 0; w >>= 1) {
                    #pragma unroll
                    for (int u = 0; u < w; ++u)
                        acc[u] = R::apply(acc[u], acc[u + w]);
                }
                const V total = warp_reduce<V, Op>(acc[0]);
                __syncwarp(); // every lane has read the slot before lane 0 lets the producer refill it
                if (lane == 0) {
                    // (SEG: a tile with a block start inside publishes its tail as a prefix)
                    const uint32_t g = tile & gmask;
                    const bool first = SEG ? tile == 0 || (seg_bits & 1u) : ((tile + p.tile_off) & p.seg_mask) == 0;
                    if constexpr (SEG)
                        s_mseg[m] = seg_bits;
                    s_local[m] = g ? gacc : R::identity();
                    gacc = g ? R::apply(gacc, total) : total;
                    if (first) {
                        V P = R::identity();
                        if (p.carry_in)
                            P = *(const V *) p.carry_in;
                        Desc<V>::publish(p.desc, tile, DESC_PREFIX, R::apply(P, total));
                    } else {
                        Desc<V>::publish(p.desc, tile, DESC_AGGREGATE, total);
                    }
                    s_mtile[m] = tile;
                    s_total[m] = total;
                    mbar_arrive(&s_agg[m]);
                    DBG_ADD(8, DBG_CLOCK() - c1);
                }
                __syncwarp();
                if (p.cyc.world) {
                    // block-cyclic multi-GPU scan: the LAST tile of a block sends the block total
                    // to every rank (one 16-byte store each).  This warp resolves the tile's prefix
                    // within the block itself -- it never waits for another GPU, whereas the
                    // look-back warps below do (block offsets): totals must not queue behind them.
                    const uint32_t lbt = p.cyc.log2_block_tiles;
                    if ((tile & ((1u << lbt) - 1u)) == (1u << lbt) - 1u) {
                        const V rel = lbt ? look_back(tile) : R::identity();
                        uint64_t bits = 0;
                        const V agg = R::apply(rel, total);
                        memcpy(&bits, &agg, sizeof(V));
                        if (lane < p.cyc.world)
                            st_entry_sys(p.cyc.table[lane] + 2 * ((size_t) (tile >> lbt) * p.cyc.table_stride + p.cyc.rank),
                                         bits, p.cyc.epoch);
                    }
                }
            }
            if (lane == 0)
                DBG_FLUSH();
            return;
        }
        // Look-back warps (LB of them, taking the CTA's tiles in turn): resolve
        // the exclusive prefix of a tile by decoupled look-back over windows of
        // 32 descriptors, publish the inclusive prefix and hand the exclusive
        // one to the compute warps.
        if (warp > WARPS) {
            V GP = R::identity(); // exclusive prefix of the current group (groups need LB == 1)
            for (uint32_t k = warp - (WARPS + 1);; k += LB) {
                const uint32_t m = k % M;
                long long c0 = DBG_CLOCK();
                mbar_wait(&s_agg[m], (k / M) & 1);
                const uint32_t tile = s_mtile[m];
                if (tile == NONE)
                    break;
                long long c1 = DBG_CLOCK();
                DBG_ADD(7, c1 - c0);
                DBG_ADD(5, 1);
                const V total = s_total[m];
                // SEG: no prefix to find when the tile begins with a block start; nothing to
                // publish when it contains one (the aggregate warp has done that already)
                const uint32_t g = tile & gmask;
                const bool first = SEG ? tile == 0 || (s_mseg[m] & 2u) : ((tile + p.tile_off) & p.seg_mask) == 0;
                V P = R::identity();
                if (p.seeds) {
                    // SEEDED: the caller knows every tile's exclusive prefix (tile sums
                    // from a reduce pass it had to make anyway) -- no chain, no polling
                    P = ((const V *) p.seeds)[p.reverse ? p.ntiles - 1 - tile : tile];
                } else {
                    if (g == 0) {
                        // only the group's first tile looks back through global memory ...
                        GP = R::identity();
                        if (!first)
                            GP = look_back(tile);
                        else if (p.carry_in)
                            GP = *(const V *) p.carry_in;
                    }
                    // ... the others continue from the aggregates this CTA holds
                    P = g ? R::apply(GP, s_local[m]) : GP;
                    if (!first && lane == 0 && !(SEG && (s_mseg[m] & 1u)))
                        Desc<V>::publish(p.desc, tile, DESC_PREFIX, R::apply(P, total));
                }
                if (p.cyc.world) {
                    // block-cyclic multi-GPU scan: P is relative to the block so far
                    const uint32_t lbt = p.cyc.log2_block_tiles, W = p.cyc.world;
                    const uint32_t j = tile >> lbt, i = tile & ((1u << lbt) - 1u);
                    bool ok = true;
                    // block offsets accumulate in double for float data (up to 4096 block totals
                    // are chained: a float chain would lose the 1e-5 the single-GPU scan keeps)
                    using A = typename std::conditional<std::is_same<V, float>::value, double, V>::type;
                    using RA = Red<A, Op>;
                    auto poll_blockoff = [&](uint32_t jj, uint64_t &bits) {
                        uint64_t e = 0, t0 = 0;
                        uint32_t spins = 0;
                        while (true) {
                            ld_relaxed_b128(p.cyc.blockoff + 2 * (size_t) jj, bits, e);
                            if (e == p.cyc.epoch)
                                return true;
                            if ((++spins & 255u) == 0) {
                                uint64_t now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                if (t0 == 0)
                                    t0 = now;
                                else if (now - t0 > 21ull * 1000 * 1000 * 1000)
                                    return false;
                            }
                            __nanosleep(100);
                        }
                    };
                    A off = RA::identity();
                    if (i == 0) {
                        // first tile: offset of block (j, rank) = offset of this rank's previous
                        // block (+) the totals of the W blocks between the two in global order:
                        // (j - 1, rank .. W - 1), then (j, 0 .. rank - 1) -- a fixed order
                        uint64_t bits = 0;
                        if (j > 0 && lane == 0) {
                            ok &= poll_blockoff(j - 1, bits);
                            memcpy(&off, &bits, sizeof(A));
                        }
                        const bool prev = lane >= p.cyc.rank; // this lane's entry belongs to round j - 1
                        V v = R::identity();
                        if (lane < W && (!prev || j > 0)) {
                            const uint64_t *tab = p.cyc.table[p.cyc.rank];
                            ok &= wait_entry_sys(tab + 2 * ((size_t) (prev ? j - 1 : j) * p.cyc.table_stride + lane),
                                                 p.cyc.epoch, bits);
                            memcpy(&v, &bits, sizeof(V));
                        }
                        if (j > 0) {
                            for (uint32_t r = p.cyc.rank; r < W; ++r)
                                off = RA::apply(off, (A) shfl_idx(v, (int) r));
                        }
                        for (uint32_t r = 0; r < p.cyc.rank; ++r)
                            off = RA::apply(off, (A) shfl_idx(v, (int) r));
                        if (lane == 0) {
                            bits = 0;
                            memcpy(&bits, &off, sizeof(A));
                            st_relaxed_b128(p.cyc.blockoff + 2 * (size_t) j, bits, p.cyc.epoch);
                        }
                    } else {
                        uint64_t bits = 0;
                        if (lane == 0)
                            ok &= poll_blockoff(j, bits);
                        bits = __shfl_sync(FULL_MASK, bits, 0);
                        memcpy(&off, &bits, sizeof(A));
                    }
                    if (!__all_sync(FULL_MASK, ok) && lane == 0)
                        *p.cyc.error = 1;
                    P = (V) RA::apply(off, (A) P);
                }
                if (lane == 0) {
                    s_prefix[m] = P;
                    if (p.carry_out && !p.seeds && tile == p.ntiles - 1)
                        *(V *) p.carry_out = R::apply(P, total);
                    mbar_arrive(&s_pref[m]);
                    DBG_ADD(2, DBG_CLOCK() - c1);
                }
                __syncwarp();
            }
            if (lane == 0)
                DBG_FLUSH();
            return;
        }
    }

    // ---- compute warps
    // SEG: position inside a block of this thread's vectors when the tile starts a block
    uint32_t seg_r[SEG ? J : 1] = {};
    if constexpr (SEG) {
        #pragma unroll
        for (int j = 0; j < J; ++j)
            seg_r[j] = ((warp * (J * 32) + j * 32 + lane) * N) % p.bs;
    }
    long long loop0 = DBG_CLOCK();
    for (uint32_t k = 0;; ++k) {
        const uint32_t slot = k % S;
        const uint32_t parity = (k / S) & 1;
        // ticket of the tile that will refill this slot: requested now, needed
        // only after the tile has been read (hides the atomic's round trip)
        uint32_t next_ticket = 0;
        if (tid == 0)
            next_ticket = next_tile();
        long long w0 = DBG_CLOCK();
        mbar_wait(&s_full[slot], parity);
        if (tid == 32)
            DBG_ADD(4, DBG_CLOCK() - w0);
        const uint32_t tile = s_tile[slot];
        if (tile == NONE)
            break;
        const uint32_t ptile = p.reverse ? p.ntiles - 1 - tile : tile;
        const uint64_t base = (uint64_t) ptile * TILE;
        const bool bulk = base + TILE <= p.size; // CTA-uniform
        const uint4 *slot_ptr = slots + (size_t) slot * VECS;
        uint32_t sv[SEG ? J : 1] = {};
        if constexpr (SEG) {
            const uint32_t spos = s_spos[slot];
            #pragma unroll
            for (int j = 0; j < J; ++j)
                sv[j] = seg_advance(spos, seg_r[j], p.bs);
        }

        // ---- load (logical vector lvi of the tile <-> physical vector pvi)
        V e[J][N];
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t lvi = warp * (J * 32) + j * 32 + lane;
            const uint32_t pvi = p.reverse ? VECS - 1 - lvi : lvi;
            const uint64_t pb = base + (uint64_t) pvi * N;
            if (bulk || pb + N <= p.size) {
                Vec16<T> v;
                if (bulk)
                    v.raw = slot_ptr[pvi];
                else
                    v.raw = ld_stream_coherent(in + pb);
                if (p.reverse) {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        e[j][kk] = to_value<T>(v.elem[N - 1 - kk]);
                } else {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        e[j][kk] = to_value<T>(v.elem[kk]);
                }
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk) {
                    const uint64_t pi = pb + (p.reverse ? N - 1 - kk : kk);
                    e[j][kk] = pi < p.size ? to_value<T>(in[pi]) : R::identity();
                }
            }
        }

        // ---- the tile now lives in registers (and the helper warps are done
        // with it: they run ahead): hand the slot back to the producer, which
        // refills it with the tile S steps ahead
        compute_sync<THREADS>();
        if (tid == 0) {
            long long i0 = DBG_CLOCK();
            if constexpr (CHAIN)
                mbar_wait(&s_agg[k % M], (k / M) & 1); // the aggregate warp has read the slot
            long long i1 = DBG_CLOCK();
            issue(slot, next_ticket);
            DBG_ADD(11, i1 - i0);
            DBG_ADD(9, DBG_CLOCK() - i1);
        }

        // ---- tile-local scan
        V carry[J];
        const bool seg_one = SEG && p.bs >= (uint32_t) N; // at most one block start per vector (CTA-uniform)
        if constexpr (SEG) {
            if (seg_one)
                TileScanSeg<T, Op, J, THREADS, true>::run(e, carry, sv, p.bs, lane, warp, s_warp, s_wt);
            else
                TileScanSeg<T, Op, J, THREADS, false>::run(e, carry, sv, p.bs, lane, warp, s_warp, s_wt);
        } else
            TileScan<T, Op, J, THREADS, CHAIN>::run(e, carry, p.log2_bs, lane, warp, s_warp);

        // ---- prefix of the tile (resolved ahead of time by the look-back warps)
        if constexpr (CHAIN) {
            long long w1 = DBG_CLOCK();
            mbar_wait(&s_pref[k % M], (k / M) & 1);
            if (tid == 32)
                DBG_ADD(3, DBG_CLOCK() - w1);
            const V P = s_prefix[k % M];
            #pragma unroll
            for (int j = 0; j < J; ++j) {
                // (SEG: only when the vector's block starts in front of the tile)
                if (!SEG || sv[j] > (warp * (J * 32) + j * 32 + lane) * N)
                    carry[j] = R::apply(P, carry[j]);
            }
        }

        // ---- results: 128-bit streaming stores straight from registers
        if (p.debug & 1)
            compute_sync<THREADS>();
        const uint32_t mask = CHAIN ? 0xffffffffu : ((1u << p.log2_bs) - 1u);
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t lvi = warp * (J * 32) + j * 32 + lane;
            const uint32_t pvi = p.reverse ? VECS - 1 - lvi : lvi;
            const uint64_t pb = base + (uint64_t) pvi * N;
            V res[N];
            if constexpr (SEG) {
                if (seg_one)
                    TileScanSeg<T, Op, J, THREADS, true>::results(e[j], carry[j], sv[j], p.bs, p.exclusive, res);
                else
                    TileScanSeg<T, Op, J, THREADS, false>::results(e[j], carry[j], sv[j], p.bs, p.exclusive, res);
            } else if (p.exclusive) {
                res[0] = carry[j];
                #pragma unroll
                for (int kk = 1; kk < N; ++kk)
                    res[kk] = (kk & mask) ? R::apply(carry[j], e[j][kk - 1]) : R::identity();
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk)
                    res[kk] = R::apply(carry[j], e[j][kk]);
            }
            if (bulk || pb + N <= p.size) {
                Vec16<T> v;
                if (p.reverse) {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        v.elem[N - 1 - kk] = from_value<T>(res[kk]);
                } else {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        v.elem[kk] = from_value<T>(res[kk]);
                }
                if (p.debug & 4)
                    *(uint4 *) (out + pb) = v.raw;
                else
                    st_stream(out + pb, v.raw);
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk) {
                    const uint64_t pi = pb + (p.reverse ? N - 1 - kk : kk);
                    if (pi < p.size)
                        out[pi] = from_value<T>(res[kk]);
                }
            }
        }
    }
    if (tid == 0 && p.done) {
        // per-stream counters (no memset per call): the CTA that finishes last leaves them at zero
        __threadfence();
        if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
            *p.ticket = 0;
            *p.done = 0;
        }
    }
    if (tid == 32)
        DBG_ADD(10, DBG_CLOCK() - loop0);
    if (tid == 0 || tid == 32)
        DBG_FLUSH();
}


/// Bounded spinning for the waits of scan_ahead_kernel: traps (a sticky error, i.e. a loud failure) when a
/// descriptor has not appeared 20 s after the first 1024 polls
struct SpinGuard {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    B200_DEVICE void tick() {
        if ((++spins & 1023u) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > 20ull * 1000 * 1000 * 1000)
                __trap();
        }
    }
};

// ---------------------------------------------------------------- AHEAD: the chained scan as two streams
//
// What bounds the chained kernel above is not a throughput but a latency: the prefix of a tile exists ~3 us
// after the tile has landed (the tiles in front of it land at about the same time, have to be reduced and
// published, and the prefix front hops over them one L2 round trip at a time), and the ring plus the
// registers of a CTA cannot hold three more microseconds of the stream.  Here the reduction of a tile is
// taken out of that chain.  Three streams, all fed by tickets (tiles and roles go to CTAs in the order in
// which they ask, so whatever a CTA waits for is owned by a CTA that is already running):
//   REDUCE  one warp per CTA, its own ticket counter and a ring of half-tile bulk copies, runs `lag` tiles
//           (tens of MB) AHEAD of the scan and publishes tile aggregates.  It throttles itself against the
//           scan's ticket counter so that its read-ahead stays inside the 126 MB L2 -- without ever blocking.
//   PREFIX  one warp of the first CTA to start: the running reduction over the aggregates, in tile order,
//           256 descriptors per batch.  It turns every AGGREGATE descriptor into an (inclusive) PREFIX: no
//           tile walks over its predecessors (the look-back of the kernel above), summation order is fixed.
//   SCAN    the streaming kernel with a ring of two slots.  Its tiles were read from HBM moments ago (the second
//           read is an L2 hit: HBM traffic stays at 8 B per element), and its prefix warp only waits for the
//           descriptor in front of a tile to become a PREFIX -- from the moment the tile's ticket is DRAWN, one
//           step before its load is requested, three before the compute warps reach it.
// Correctness does not depend on the lag: every wait is on a descriptor that some running CTA will publish.
template <typename T, int Op, int J, int S, int RS, int THREADS>
__global__ void __launch_bounds__(THREADS + 64, 2)
scan_ahead_kernel(const FastParams p) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    constexpr int N = VecInfo<T>::N;
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t VECS = THREADS * J;
    constexpr uint32_t TILE = VECS * N;
    constexpr uint32_t TILE_BYTES = VECS * 16;
    constexpr uint32_t UPT = 2;                   // reduce units per tile
    constexpr uint32_t RVECS = VECS / UPT, RBYTES = TILE_BYTES / UPT, RELEMS = TILE / UPT;
    constexpr uint32_t NONE = 0xffffffffu;
    constexpr int M = S + 2;                      // tickets are drawn S + 1 entries ahead of the compute warps

    extern __shared__ __align__(128) uint8_t sf_smem[];
    uint4 *slots = (uint4 *) sf_smem;             // S * VECS vectors: scan ring
    uint4 *rslots = slots + (size_t) S * VECS;    // RS * RVECS vectors: reduce ring
    __shared__ __align__(8) uint64_t s_full[S];   // scan ring: bulk load has landed
    __shared__ __align__(8) uint64_t s_rfull[RS]; // reduce ring: bulk load has landed
    __shared__ __align__(8) uint64_t s_iss[M];    // s_qtile entry is valid
    __shared__ __align__(8) uint64_t s_pref[M];   // s_prefix entry is valid
    __shared__ uint32_t s_qtile[M];
    __shared__ uint32_t s_rtile[RS];
    __shared__ uint32_t s_role;
    __shared__ V s_prefix[M];
    __shared__ V s_warp[WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T *in = (const T *) p.in;
    T *out = (T *) p.out;
    uint32_t *scan_ticket = p.ticket, *reduce_ticket = p.ticket + 1;

    auto issue = [&](uint32_t slot, uint32_t t) { // (thread 0)
        if (t == NONE) {
            mbar_arrive(&s_full[slot]);
            return;
        }
        const uint32_t pt = p.reverse ? p.ntiles - 1 - t : t;
        const uint64_t base = (uint64_t) pt * TILE;
        if (base + TILE <= p.size) {
            mbar_arrive_expect_tx(&s_full[slot], TILE_BYTES);
            bulk_g2s(slots + (size_t) slot * VECS, in + base, TILE_BYTES, &s_full[slot]);
        } else {
            mbar_arrive(&s_full[slot]); // partial tile: loaded directly by its readers
        }
    };
    auto draw = [&](uint32_t entry, uint32_t t) { // (thread 0)
        s_qtile[entry % M] = t < p.ntiles ? t : NONE;
        mbar_arrive(&s_iss[entry % M]);
    };

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < S; ++s)
            mbar_init(&s_full[s], 1);
        #pragma unroll
        for (int s = 0; s < RS; ++s)
            mbar_init(&s_rfull[s], 1);
        #pragma unroll
        for (int s = 0; s < M; ++s) {
            mbar_init(&s_iss[s], 1);
            mbar_init(&s_pref[s], 1);
        }
        mbar_fence_init();
        // Roles are handed out in the order in which CTAs START: the first one runs the prefix stream.  Every
        // stream a CTA may wait for is therefore owned by a CTA that is already running (the reduce and scan
        // streams through their tickets, the prefix stream through this one) -- forward progress does not
        // depend on the whole grid being resident.
        s_role = atomicAdd(p.ticket + 2, 1u);
        if (s_role != 0) {
            for (uint32_t e = 0; e <= (uint32_t) S; ++e)
                draw(e, atomicAdd(scan_ticket, 1u));
            for (uint32_t e = 0; e < (uint32_t) S; ++e)
                issue(e, s_qtile[e]);
        }
    }
    __syncthreads();

    // ---- PREFIX stream (the first CTA to start, one warp): the running reduction over the tile aggregates, in
    // tile order, batches of 32 * PW descriptors with the next batch's loads in flight -- one sequential
    // stream instead of a look-back per tile (no tile ever walks over its predecessors; the front of known
    // prefixes moves 32 * PW tiles per L2 round trip and stays far in front of the scan stream).  It turns
    // every AGGREGATE entry into the PREFIX entry (inclusive) that the prefix warps of the other CTAs wait for.
    if (s_role == 0) {
        if (warp != 0)
            return;
        constexpr int PW = 8;
        V val[PW], nval[PW];
        uint32_t st[PW], nst[PW];
        auto fetch = [&](uint32_t base, V (&v)[PW], uint32_t (&s)[PW]) {
            #pragma unroll
            for (int w = 0; w < PW; ++w) {
                const uint32_t idx = base + 32 * w + lane;
                v[w] = R::identity();
                s[w] = DESC_AGGREGATE;
                if (idx < p.ntiles)
                    s[w] = Desc<V>::observe(p.desc, idx, v[w]);
            }
        };
        // float prefixes are carried in double: up to 2^19 tile aggregates are chained here, and a float chain
        // would spend a good part of the 1e-5 the scan promises (as the block-cyclic offsets above)
        using A = typename std::conditional<std::is_same<V, float>::value, double, V>::type;
        using RA = Red<A, Op>;
        A carry = RA::identity();
        fetch(0, val, st);
        for (uint32_t base = 0; base < p.ntiles; base += 32 * PW) {
            // all entries of the batch must be there (the reduce stream is usually far ahead)
            SpinGuard guard;
            while (true) {
                bool missing = false;
                #pragma unroll
                for (int w = 0; w < PW; ++w)
                    missing |= st[w] == DESC_INVALID;
                if (!__any_sync(FULL_MASK, missing))
                    break;
                __nanosleep(200);
                guard.tick();
                #pragma unroll
                for (int w = 0; w < PW; ++w) {
                    const uint32_t idx = base + 32 * w + lane;
                    if (st[w] == DESC_INVALID)
                        st[w] = Desc<V>::observe(p.desc, idx, val[w]);
                }
            }
            if (base + 32 * PW < p.ntiles)
                fetch(base + 32 * PW, nval, nst);
            #pragma unroll
            for (int w = 0; w < PW; ++w) {
                const uint32_t idx = base + 32 * w + lane;
                // block starts (and tile 0) hold their inclusive prefix already
                const bool head = ((idx + p.tile_off) & p.seg_mask) == 0 || idx == 0;
                A v = (A) val[w];
                bool f = head;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const A uv = shfl_up(v, d);
                    const bool uf = __shfl_up_sync(FULL_MASK, (int) f, d) != 0;
                    if (lane >= (uint32_t) d) {
                        if (!f)
                            v = RA::apply(uv, v);
                        f = f || uf;
                    }
                }
                if (!f)
                    v = RA::apply(carry, v);
                carry = shfl_idx(v, 31);
                if (idx < p.ntiles) {
                    // (block starts were published as prefixes by the reduce stream)
                    if (((idx + p.tile_off) & p.seg_mask) != 0)
                        Desc<V>::publish(p.desc, idx, DESC_PREFIX, (V) v);
                    if (p.carry_out && idx == p.ntiles - 1)
                        *(V *) p.carry_out = (V) v;
                }
            }
            #pragma unroll
            for (int w = 0; w < PW; ++w) {
                val[w] = nval[w];
                st[w] = nst[w];
            }
        }
        return;
    }

    // ---- REDUCE stream (one warp): tile aggregates, `lag` tiles ahead of the scan stream
    if (warp == WARPS) {
        uint32_t q_issue = 0, q_proc = 0;  // units requested / reduced
        uint32_t cur = 0;                  // tile of the unit requested next
        uint32_t seen = 0, seen_next = 0;  // (lane 0) scan ticket counter when last looked at (read ahead as well)
        // (lane 0) tickets are drawn one tile ahead of their use: the round trip of the atomic (~1 us under
        // load) must not stand between two reductions
        uint32_t next_r = lane == 0 ? atomicAdd(reduce_ticket, 1u) : 0u;
        bool finished = false;
        constexpr int AU = 8;
        static_assert(RVECS % (32 * AU) == 0, "a unit must be a multiple of 256 vectors");
        V acc[AU];
        while (true) {
            // -- keep the ring full
            while (!finished && q_issue < q_proc + RS) {
                if (q_issue % UPT == 0) {
                    const uint32_t r = __shfl_sync(FULL_MASK, next_r, 0);
                    if (r >= p.ntiles) {
                        finished = true;
                        break;
                    }
                    // read-ahead stays inside L2: a tile is requested only once the scan stream is within `lag`
                    // tiles of it.  Never blocking: what has landed is reduced first, then the warp looks again
                    // (the held ticket lies far in front of every tile the scan stream is waiting for).
                    bool ok = true;
                    if (lane == 0) {
                        ok = r < seen + p.lag;
                        if (!ok) {
                            seen = max(seen, seen_next);
                            ok = r < seen + p.lag;
                        }
                        if (!ok) {
                            seen = *(volatile uint32_t *) scan_ticket;
                            ok = r < seen + p.lag;
                        }
                    }
                    if (!__shfl_sync(FULL_MASK, (int) ok, 0))
                        break;
                    cur = r;
                    if (lane == 0) {
                        next_r = atomicAdd(reduce_ticket, 1u);
                        seen_next = *(volatile uint32_t *) scan_ticket;
                    }
                }
                if (lane == 0) {
                    const uint32_t slot = q_issue % RS, h = q_issue % UPT;
                    s_rtile[slot] = cur;
                    const uint32_t pt = p.reverse ? p.ntiles - 1 - cur : cur;
                    const uint64_t base = (uint64_t) pt * TILE;
                    if (base + TILE <= p.size) {
                        mbar_arrive_expect_tx(&s_rfull[slot], RBYTES);
                        bulk_g2s(rslots + (size_t) slot * RVECS, in + base + (uint64_t) h * RELEMS, RBYTES, &s_rfull[slot]);
                    } else {
                        mbar_arrive(&s_rfull[slot]); // partial tile: direct loads
                    }
                }
                ++q_issue;
            }
            if (q_proc == q_issue) {
                if (finished)
                    break;
                __nanosleep(500);
                continue;
            }
            // -- reduce one unit
            const uint32_t slot = q_proc % RS, h = q_proc % UPT;
            mbar_wait(&s_rfull[slot], (q_proc / RS) & 1);
            const uint32_t tile = s_rtile[slot];
            const uint32_t ptile = p.reverse ? p.ntiles - 1 - tile : tile;
            const uint64_t base = (uint64_t) ptile * TILE;
            if (h == 0) {
                #pragma unroll
                for (int u = 0; u < AU; ++u)
                    acc[u] = R::identity();
            }
            if (base + TILE <= p.size) {
                const uint4 *slot_ptr = rslots + (size_t) slot * RVECS;
                #pragma unroll 1
                for (uint32_t i = lane; i < RVECS; i += 32 * AU) {
                    Vec16<T> v[AU];
                    #pragma unroll
                    for (int u = 0; u < AU; ++u)
                        v[u].raw = slot_ptr[i + u * 32];
                    #pragma unroll
                    for (int u = 0; u < AU; ++u) {
                        V t[N];
                        #pragma unroll
                        for (int kk = 0; kk < N; ++kk)
                            t[kk] = to_value<T>(v[u].elem[kk]);
                        #pragma unroll
                        for (int w = N / 2; w > 0; w >>= 1) {
                            #pragma unroll
                            for (int kk = 0; kk < w; ++kk)
                                t[kk] = R::apply(t[kk], t[kk + w]);
                        }
                        acc[u] = R::apply(acc[u], t[0]);
                    }
                }
            } else {
                const uint64_t lo = base + (uint64_t) h * RELEMS, hi = min(lo + RELEMS, p.size);
                for (uint64_t i = lo + lane; i < hi; i += 32)
                    acc[0] = R::apply(acc[0], to_value<T>(in[i]));
            }
            __syncwarp();
            ++q_proc;
            if (h == UPT - 1) {
                V a[AU];
                #pragma unroll
                for (int u = 0; u < AU; ++u)
                    a[u] = acc[u];
                #pragma unroll
                for (int w = AU / 2; w > 0; w >>= 1) {
                    #pragma unroll
                    for (int u = 0; u < w; ++u)
                        a[u] = R::apply(a[u], a[u + w]);
                }
                const V total = warp_reduce<V, Op>(a[0]);
                if (lane == 0) {
                    if (((tile + p.tile_off) & p.seg_mask) == 0) {
                        V P = R::identity();
                        if (p.carry_in)
                            P = *(const V *) p.carry_in;
                        Desc<V>::publish(p.desc, tile, DESC_PREFIX, R::apply(P, total));
                    } else {
                        Desc<V>::publish(p.desc, tile, DESC_AGGREGATE, total);
                    }
                }
            }
        }
        return;
    }

    // ---- prefix warp: waits for the inclusive prefix of the tile in front (written by the prefix stream),
    // from the moment a ticket is drawn
    if (warp == WARPS + 1) {
        for (uint32_t k = 0;; ++k) {
            const uint32_t m = k % M;
            mbar_wait(&s_iss[m], (k / M) & 1);
            const uint32_t tile = s_qtile[m];
            if (tile == NONE)
                break;
            V P = R::identity();
            if (lane == 0) {
                if (((tile + p.tile_off) & p.seg_mask) == 0) {
                    if (p.carry_in)
                        P = *(const V *) p.carry_in;
                } else if (tile > 0) {
                    SpinGuard guard; // (a lost prefix must fail loudly, not hang)
                    while (Desc<V>::observe(p.desc, tile - 1, P) != DESC_PREFIX) {
                        __nanosleep(100);
                        guard.tick();
                    }
                }
            } else if (lane == 1 && p.in == p.out) {
                // in place: the tile must not be overwritten before the reduce stream has read it (nothing else
                // orders the two -- the scan of a tile needs the aggregates in front of it, not its own)
                V own;
                SpinGuard guard;
                while (Desc<V>::observe(p.desc, tile, own) == DESC_INVALID) {
                    __nanosleep(100);
                    guard.tick();
                }
            }
            __syncwarp();
            if (lane == 0) {
                s_prefix[m] = P;
                mbar_arrive(&s_pref[m]);
            }
            __syncwarp();
        }
        return;
    }

    // ---- compute warps
    for (uint32_t k = 0;; ++k) {
        const uint32_t slot = k % S;
        // the ticket of the entry S + 1 steps ahead: requested now, handed on after the tile has been read
        uint32_t drawn = 0;
        if (tid == 0)
            drawn = atomicAdd(scan_ticket, 1u);
        mbar_wait(&s_full[slot], (k / S) & 1);
        const uint32_t tile = s_qtile[k % M];
        if (tile == NONE)
            break;
        const uint32_t ptile = p.reverse ? p.ntiles - 1 - tile : tile;
        const uint64_t base = (uint64_t) ptile * TILE;
        const bool bulk = base + TILE <= p.size; // CTA-uniform
        const uint4 *slot_ptr = slots + (size_t) slot * VECS;

        V e[J][N];
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t lvi = warp * (J * 32) + j * 32 + lane;
            const uint32_t pvi = p.reverse ? VECS - 1 - lvi : lvi;
            const uint64_t pb = base + (uint64_t) pvi * N;
            if (bulk || pb + N <= p.size) {
                Vec16<T> v;
                if (bulk)
                    v.raw = slot_ptr[pvi];
                else
                    v.raw = ld_stream_coherent(in + pb);
                if (p.reverse) {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        e[j][kk] = to_value<T>(v.elem[N - 1 - kk]);
                } else {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        e[j][kk] = to_value<T>(v.elem[kk]);
                }
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk) {
                    const uint64_t pi = pb + (p.reverse ? N - 1 - kk : kk);
                    e[j][kk] = pi < p.size ? to_value<T>(in[pi]) : R::identity();
                }
            }
        }

        // the tile lives in registers: refill the slot, pass the new ticket on to the look-back warp
        compute_sync<THREADS>();
        if (tid == 0) {
            draw(k + S + 1, drawn);
            issue(slot, s_qtile[(k + S) % M]);
        }

        V carry[J];
        TileScan<T, Op, J, THREADS, true>::run(e, carry, 0, lane, warp, s_warp);

        mbar_wait(&s_pref[k % M], (k / M) & 1);
        const V P = s_prefix[k % M];
        #pragma unroll
        for (int j = 0; j < J; ++j)
            carry[j] = R::apply(P, carry[j]);

        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t lvi = warp * (J * 32) + j * 32 + lane;
            const uint32_t pvi = p.reverse ? VECS - 1 - lvi : lvi;
            const uint64_t pb = base + (uint64_t) pvi * N;
            V res[N];
            if (p.exclusive) {
                res[0] = carry[j];
                #pragma unroll
                for (int kk = 1; kk < N; ++kk)
                    res[kk] = R::apply(carry[j], e[j][kk - 1]);
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk)
                    res[kk] = R::apply(carry[j], e[j][kk]);
            }
            if (bulk || pb + N <= p.size) {
                Vec16<T> v;
                if (p.reverse) {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        v.elem[N - 1 - kk] = from_value<T>(res[kk]);
                } else {
                    #pragma unroll
                    for (int kk = 0; kk < N; ++kk)
                        v.elem[kk] = from_value<T>(res[kk]);
                }
                st_stream(out + pb, v.raw);
            } else {
                #pragma unroll
                for (int kk = 0; kk < N; ++kk) {
                    const uint64_t pi = pb + (p.reverse ? N - 1 - kk : kk);
                    if (pi < p.size)
                        out[pi] = from_value<T>(res[kk]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- dispatch

/// 0: general kernel only, otherwise (default) the streaming kernels.
/// Development switch.
static int scan_path() {
    static int path = -1;
    if (path < 0) {
        const char *s = getenv("B200_SCAN_PATH");
        path = s ? atoi(s) : 2;
    }
    return path;
}

/// Cap of log2(tiles per ticket) of the chained scan over blocks of several tiles (development switch
/// B200_SCAN_GROUP)
static uint32_t scan_group_log2() {
    static int lg = -1;
    if (lg < 0) {
        const char *s = getenv("B200_SCAN_GROUP");
        lg = s ? atoi(s) : 4;
        if (lg < 0 || lg > 8)
            lg = 4;
    }
    return (uint32_t) lg;
}

template <typename K> static int prepare_kernel(K kernel, int threads, size_t smem, int *occupancy) {
    if (smem > 48 * 1024)
        B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) smem));
    int occ = 0;
    B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    *occupancy = occ < 1 ? 1 : occ;
    if (getenv("B200_DEBUG"))
        fprintf(stderr, "b200: scan_stream_kernel: %d threads, %d CTAs/SM, %zu B dynamic smem\n",
                threads, occ, smem);
    return B200_OK;
}

// Tile geometry: THREADS compute threads x J vectors per tile, S ring slots, LB
// look-back warps.  Default 512 x 4 = 32 KiB tiles, 3 slots, two CTAs per SM:
// measured on B200 at 2^28 fp32, POW2 0.318 ms (6.7 TB/s) and CHAIN 0.38 ms
// (5.6 TB/s).  256 x 4 with 6 slots: POW2 the same, CHAIN 0.44 ms; a second
// look-back warp makes CHAIN slower (more polling, same dependency latency).
template <int THREADS_, int J_, int S_, int LB_ = 2, int LBW_ = 1> struct Geom {
    static constexpr int THREADS = THREADS_, J = J_, S = S_, LB = LB_, LBW = LBW_;
};
using DefaultGeom = Geom<512, 4, 3, 1, B200_SCAN_LBW>;

template <typename T, int Op, bool CHAIN, typename G, bool SEG = false>
static int launch_stream(const ScanCall &c, FastParams &p) {
    using V = typename ValueOf<T>::type;
    constexpr size_t SMEM = (size_t) G::S * G::THREADS * G::J * 16;
    constexpr int BLOCK = G::THREADS + (CHAIN ? 32 + 32 * G::LB : 0);
    auto kernel = scan_stream_kernel<T, Op, G::J, G::S, G::THREADS, CHAIN, CHAIN ? G::LB : 0, CHAIN ? G::LBW : 1, SEG>;

    // per device: opt in to the dynamic shared memory size, query residency
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = dev < 64 ? occ_cache[dev].load(std::memory_order_relaxed) : 0;
    if (occ == 0) {
        int rc = prepare_kernel(kernel, BLOCK, SMEM, &occ);
        if (rc)
            return rc;
        if (dev < 64)
            occ_cache[dev].store(occ, std::memory_order_relaxed);
    }

    void *scratch = nullptr;
    unsigned int *counters = CHAIN ? nullptr : stream_ticket(c.stream);
    if (counters) {
        // independent tiles: only the ticket is needed -- the stream's own self-resetting counters
        p.ticket = counters;
        p.done = counters + 1;
    } else {
        size_t desc_bytes = CHAIN ? (size_t) p.ntiles * Desc<V>::WORDS * sizeof(uint64_t) : 0;
        scratch = temp_alloc(desc_bytes + 16, c.stream);
        if (!scratch)
            return fail(B200_ERR_CUDA, "jit_block_prefix_reduce(): out of memory");
        B200_CUDA_CHECK(cudaMemsetAsync(scratch, 0, desc_bytes + 16, c.stream));
        p.desc = (uint64_t *) scratch;
        p.ticket = (uint32_t *) ((uint8_t *) scratch + desc_bytes);
    }
    uint32_t grid = (uint32_t) std::min<uint64_t>(p.ntiles, (uint64_t) sm_count() * occ);
    kernel<<<grid, BLOCK, SMEM, c.stream>>>(p);
    if (scratch)
        temp_free(scratch, c.stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}


/// Development switches of the two-stream chained scan: B200_SCAN_AHEAD=0 turns it off, B200_SCAN_LAG sets the
/// read-ahead of the reduce stream in tiles (default 2048 tiles = 64 MiB)
static int scan_ahead_lag() {
    static int lag = -2;
    if (lag == -2) {
        const char *a = getenv("B200_SCAN_AHEAD"), *l = getenv("B200_SCAN_LAG");
        lag = (a && atoi(a) == 0) ? -1 : (l ? atoi(l) : 2048);
    }
    return lag;
}

/// Whole arrays / blocks of many tiles, at least 1024 tiles: scan_ahead_kernel.  *handled stays false when the
/// kernel does not fit twice on an SM (the chained kernel above takes over).
template <typename T, int Op, typename G> static int launch_ahead(const ScanCall &c, FastParams &p, bool *handled) {
    using V = typename ValueOf<T>::type;
    constexpr int S = 2, RS = 3;
    constexpr size_t SMEM = (size_t) (2 * S + RS) * (G::THREADS * G::J / 2) * 16;
    constexpr int BLOCK = G::THREADS + 64;
    auto kernel = scan_ahead_kernel<T, Op, G::J, S, RS, G::THREADS>;

    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = dev < 64 ? occ_cache[dev].load(std::memory_order_relaxed) : 0;
    if (occ == 0) {
        int rc = prepare_kernel(kernel, BLOCK, SMEM, &occ);
        if (rc)
            return rc;
        if (dev < 64)
            occ_cache[dev].store(occ, std::memory_order_relaxed);
    }
    if (occ < 2)
        return B200_OK;
    const uint32_t grid = (uint32_t) std::min<uint64_t>(p.ntiles, (uint64_t) sm_count() * occ);
    p.lag = std::max<uint32_t>((uint32_t) scan_ahead_lag(), 2 * grid);

    const size_t desc_bytes = (size_t) p.ntiles * Desc<V>::WORDS * sizeof(uint64_t);
    void *scratch = temp_alloc(desc_bytes + 16, c.stream);
    if (!scratch)
        return fail(B200_ERR_CUDA, "jit_block_prefix_reduce(): out of memory");
    B200_CUDA_CHECK(cudaMemsetAsync(scratch, 0, desc_bytes + 16, c.stream));
    p.desc = (uint64_t *) scratch;
    p.ticket = (uint32_t *) ((uint8_t *) scratch + desc_bytes); // scan tickets, reduce tickets, CTA roles
    kernel<<<grid, BLOCK, SMEM, c.stream>>>(p);
    temp_free(scratch, c.stream);
    B200_LAUNCH_CHECK();
    *handled = true;
    return B200_OK;
}

template <typename T, int Op, typename G> static int launch_fast_g(const ScanCall &c, bool *handled) {
    constexpr int N = VecInfo<T>::N;
    constexpr uint32_t TILE = G::THREADS * G::J * N;
    *handled = false;
    if (scan_path() == 0 || (((uintptr_t) c.in | (uintptr_t) c.out) & 15) != 0)
        return B200_OK;
    if (c.size >= 0xffffffffull - 2 * TILE)
        return B200_OK;

    FastParams p{};
    p.in = c.in;
    p.out = c.out;
    p.size = c.size;
    p.ntiles = (uint32_t) ceil_div(c.size, TILE);
    p.exclusive = c.exclusive;
    p.reverse = c.reverse;
    p.carry_in = c.carry_in;
    p.carry_out = c.carry_out;
    p.seeds = c.seeds;
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char *e = getenv("B200_SCAN_DBG");
            dbg = e ? atoi(e) : 0;
        }
        p.debug = (uint32_t) dbg;
    }

    bool chain;
    if (c.cyclic) {
        // every block of 2^log2_block_tiles tiles is a segment of the local chain
        chain = true;
        p.cyc = *c.cyclic;
        p.seg_mask = (1u << c.cyclic->log2_block_tiles) - 1u;
        if (c.reverse || c.seeds || (p.ntiles & p.seg_mask) != 0)
            return fail(B200_ERR_INVALID, "block-cyclic scan: whole blocks, forward only!");
    } else if (c.carry_api || c.bs >= c.size) {
        chain = true;
        p.seg_mask = 0xffffffffu;
    } else if (is_pow2(c.bs) && c.bs <= TILE) {
        chain = false;
        p.log2_bs = log2i(c.bs);
    } else if (is_pow2(c.bs)) {
        chain = true;
        uint32_t tps = (uint32_t) (c.bs / TILE);
        p.seg_mask = tps - 1;
        p.tile_off = c.reverse ? (tps - p.ntiles % tps) % tps : 0;
    } else if (c.bs >= 2 && c.bs <= 0xffffffffull && !c.carry_in && !c.carry_out && !c.seeds) {
        // any other block size: chained tiles with block starts at arbitrary elements
        p.bs = (uint32_t) c.bs;
        p.off = c.reverse ? (uint32_t) ((c.bs - ((uint64_t) p.ntiles * TILE) % c.bs) % c.bs) : 0u;
        p.seg_mask = 0xffffffffu;
        *handled = true;
        return launch_stream<T, Op, true, G, true>(c, p);
    } else {
        return B200_OK;
    }
    *handled = true;
    if (chain && !c.cyclic && !c.seeds && G::LB == 1 && p.seg_mask != 0xffffffffu) {
        // A block of up to 2^cap tiles is taken by ONE CTA, tile after tile (a "group"): no global look-back
        // at all (2^28 fp32, blocks of 2 / 4 / 8 / 16 tiles: 0.37-0.38 ms -> 0.34 / 0.34 / 0.35 / 0.36 ms).
        // Groups that are only a part of their block lose (a tile's predecessors then land one ring turn
        // later and every look-back waits that long; blocks of 32 tiles in groups of 16: 0.38 -> 0.45 ms, the
        // whole array in groups of 2 / 4: 0.39 -> 0.44 / 0.48 ms), so larger blocks keep single tiles.
        const uint32_t lg = log2i(p.seg_mask + 1);
        if (lg <= scan_group_log2() && (p.tile_off & p.seg_mask) == 0)
            p.log2_group = lg;
    }
    // (4-byte types: measured slower than the chained kernel for 64-bit elements, 0.51 vs 0.475 ms at 2^27)
    if constexpr (sizeof(T) == 4) {
        if (chain && !c.cyclic && !c.seeds && p.log2_group == 0 && p.ntiles >= 1024 && scan_ahead_lag() >= 0) {
            bool done = false;
            int rc = launch_ahead<T, Op, G>(c, p, &done);
            if (rc || done)
                return rc;
        }
    }
    if (chain)
        return launch_stream<T, Op, true, G>(c, p);
    return launch_stream<T, Op, false, G>(c, p);
}

template <typename T, int Op> static int launch_fast(const ScanCall &c, bool *handled) {
#if defined(B200_SCAN_TUNING)
    // development build: alternative geometries for fp32 Add, picked at run time
    if constexpr (std::is_same<T, float>::value && Op == B200_OP_ADD) {
        static int geom = -1;
        if (geom < 0) {
            const char *s = getenv("B200_SCAN_GEOM");
            geom = s ? atoi(s) : 0;
        }
        switch (geom) {
            case 1: return launch_fast_g<T, Op, Geom<256, 4, 6, 1>>(c, handled);
            case 2: return launch_fast_g<T, Op, Geom<512, 4, 3, 2>>(c, handled);
            case 3: return launch_fast_g<T, Op, Geom<512, 4, 3, 1>>(c, handled);
            case 4: return launch_fast_g<T, Op, Geom<512, 2, 6, 2>>(c, handled);
            case 5: return launch_fast_g<T, Op, Geom<256, 4, 4, 1>>(c, handled);
            case 6: return launch_fast_g<T, Op, Geom<384, 4, 4, 2>>(c, handled);
            case 7: return launch_fast_g<T, Op, Geom<384, 4, 4, 1>>(c, handled);
            case 8: return launch_fast_g<T, Op, Geom<512, 8, 2, 1>>(c, handled);
            case 9: return launch_fast_g<T, Op, Geom<512, 8, 3, 1>>(c, handled);
            case 10: return launch_fast_g<T, Op, Geom<768, 4, 3, 1>>(c, handled);
            case 11: return launch_fast_g<T, Op, Geom<896, 4, 3, 1>>(c, handled);
            case 12: return launch_fast_g<T, Op, Geom<512, 6, 3, 1>>(c, handled);
            case 13: return launch_fast_g<T, Op, Geom<512, 4, 3, 1, 2>>(c, handled);
            case 14: return launch_fast_g<T, Op, Geom<512, 4, 3, 1, 4>>(c, handled);
            case 15: return launch_fast_g<T, Op, Geom<512, 8, 3, 1, 2>>(c, handled);
            case 16: return launch_fast_g<T, Op, Geom<512, 8, 3, 1, 4>>(c, handled);
            case 17: return launch_fast_g<T, Op, Geom<512, 4, 3, 1, 8>>(c, handled);
            case 18: return launch_fast_g<T, Op, Geom<512, 4, 6, 1, 4>>(c, handled);
            default: break;
        }
    }
#endif
    return launch_fast_g<T, Op, DefaultGeom>(c, handled);
}

typedef int (*FastFn)(const ScanCall &, bool *);

template <typename T, bool Bits> static FastFn pick_fast_op(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_fast<T, B200_OP_ADD>;
        case B200_OP_MUL: return launch_fast<T, B200_OP_MUL>;
        case B200_OP_MIN: return launch_fast<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_fast<T, B200_OP_MAX>;
        case B200_OP_AND: if constexpr (Bits) return launch_fast<T, B200_OP_AND>; else return nullptr;
        case B200_OP_OR:  if constexpr (Bits) return launch_fast<T, B200_OP_OR>; else return nullptr;
        default: return nullptr;
    }
}

template <typename T> static FastFn pick_fast_minmax(int op) {
    switch (op) {
        case B200_OP_MIN: return launch_fast<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_fast<T, B200_OP_MAX>;
        default: return nullptr;
    }
}

int scan_fast_dispatch(int vt, int op, const ScanCall &call, bool *handled) {
    const bool minmax = op == B200_OP_MIN || op == B200_OP_MAX;
    FastFn fn = nullptr;
    switch (vt) {
        case B200_VT_INT32:  fn = minmax ? pick_fast_minmax<int32_t>(op) : pick_fast_op<uint32_t, true>(op); break;
        case B200_VT_UINT32: fn = pick_fast_op<uint32_t, true>(op); break;
        case B200_VT_INT64:  fn = minmax ? pick_fast_minmax<int64_t>(op) : pick_fast_op<uint64_t, true>(op); break;
        case B200_VT_UINT64: fn = pick_fast_op<uint64_t, true>(op); break;
        case B200_VT_FLOAT16: fn = pick_fast_op<__half, false>(op); break;
        case B200_VT_FLOAT32: fn = pick_fast_op<float, false>(op); break;
        case B200_VT_FLOAT64: fn = pick_fast_op<double, false>(op); break;
        default: break;
    }
    *handled = false;
    if (!fn)
        return B200_OK;
    return fn(call, handled);
}

#if defined(B200_SCAN_TUNING)
extern "C" __attribute__((visibility("default"))) void b200_scan_debug(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_scan_dbg, sizeof(g_scan_dbg));
    if (reset) {
        unsigned long long zero[16] = { 0 };
        cudaMemcpyToSymbol(g_scan_dbg, zero, sizeof(zero));
    }
}
#endif

} // namespace b200
