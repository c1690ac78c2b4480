// compress.cu -- stream compaction of a byte mask into ascending indices.
//
// Replaces CUDAThreadState::compress (src/cuda_ts.cpp:683-763) and
// resources/compress.cuh (compress_small / compress_large_init /
// compress_large).  Masks of up to 32768 entries use one single-pass kernel:
//   - 256 threads x J 16-byte vectors = 16 KiB of mask per tile (the reference:
//     2 KiB), tiles handed out by an atomic ticket and chained with decoupled
//     look-back on 64-bit {status, count} descriptors;
//   - the mask is never written: the tail and a misaligned head are handled by
//     guarded loads instead of the reference's memset of the trailer
//     (src/cuda_ts.cpp:708-710, :746-748);
//   - indices are compacted per warp row in shared memory and written with
//     contiguous (coalesced) stores instead of one predicated scattered store
//     per mask byte (compress.cuh:146-149).
// Larger masks take the bit-packed two-pass path further down.
#include "common.cuh"
#include "pipeline.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace b200 {

static constexpr int COMPRESS_THREADS = 256;
static constexpr int COMPRESS_WARPS = COMPRESS_THREADS / 32;

/// bit 7 of every byte that is non-zero
B200_DEVICE uint32_t nonzero_bytes(uint32_t w) {
    return (w | ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
}

template <int J>
__global__ void __launch_bounds__(COMPRESS_THREADS)
compress_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t size,
                uint32_t ntiles, uint64_t *desc, uint32_t *ticket, uint32_t *count_out) {
    constexpr uint32_t TILE = COMPRESS_THREADS * J * 16;
    constexpr int ENTRIES = J * COMPRESS_WARPS;
    static_assert(ENTRIES <= 32, "entry scan is done by one warp");

    __shared__ uint32_t s_incl[ENTRIES];
    __shared__ uint32_t s_excl[ENTRIES];
    __shared__ uint32_t s_tile_id;
    __shared__ uint32_t s_stage[COMPRESS_WARPS][32 * 16];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mis = (uint32_t) ((uintptr_t) in & 15);

    if (tid == 0)
        s_tile_id = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile_id;

    // ---- load: virtual byte b of the aligned array (in - mis) is item b - mis
    uint32_t nz[J][4], cnt[J];
    #pragma unroll
    for (int j = 0; j < J; ++j) {
        const uint64_t vb = (uint64_t) tile * TILE + (uint64_t) (j * COMPRESS_THREADS + tid) * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (vb >= mis && vb + 16 <= size + mis) {
            v = ld_stream(in + (vb - mis));
        } else if (vb < size + mis && vb + 16 > mis) {
            uint32_t w[4] = { 0, 0, 0, 0 };
            #pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint64_t b = vb + k;
                if (b >= mis && b < size + mis)
                    w[k >> 2] |= (uint32_t) in[b - mis] << (8 * (k & 3));
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        nz[j][0] = nonzero_bytes(v.x);
        nz[j][1] = nonzero_bytes(v.y);
        nz[j][2] = nonzero_bytes(v.z);
        nz[j][3] = nonzero_bytes(v.w);
        cnt[j] = __popc(nz[j][0]) + __popc(nz[j][1]) + __popc(nz[j][2]) + __popc(nz[j][3]);
    }

    // ---- warp-level inclusive scan of the per-vector counts
    uint32_t incl[J];
    #pragma unroll
    for (int j = 0; j < J; ++j)
        incl[j] = cnt[j];
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            uint32_t up = __shfl_up_sync(FULL_MASK, incl[j], d);
            if (lane >= d)
                incl[j] += up;
        }
    }
    if (lane == 31) {
        #pragma unroll
        for (int j = 0; j < J; ++j)
            s_incl[j * COMPRESS_WARPS + warp] = incl[j];
    }
    __syncthreads();

    // ---- one warp: scan the row totals, chain with the preceding tiles
    if (warp == 0) {
        uint32_t a = lane < ENTRIES ? s_incl[lane] : 0;
        #pragma unroll
        for (int d = 1; d < ENTRIES; d <<= 1) {
            uint32_t up = __shfl_up_sync(FULL_MASK, a, d);
            if (lane >= d)
                a += up;
        }
        const uint32_t total = __shfl_sync(FULL_MASK, a, ENTRIES - 1);
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0)
                Desc<uint32_t>::publish(desc, 0, DESC_PREFIX, total);
        } else {
            if (lane == 0)
                Desc<uint32_t>::publish(desc, tile, DESC_AGGREGATE, total);
            int64_t base = (int64_t) tile - 1;
            while (true) {
                int64_t idx = base - lane;
                uint32_t val = 0, st = DESC_PREFIX;
                do {
                    if (idx >= 0)
                        st = Desc<uint32_t>::observe(desc, (uint32_t) idx, val);
                } while (__any_sync(FULL_MASK, st == DESC_INVALID));
                uint32_t ballot = __ballot_sync(FULL_MASK, st == DESC_PREFIX);
                if (ballot) {
                    uint32_t first = __ffs(ballot) - 1;
                    prefix += __reduce_add_sync(FULL_MASK, lane <= first ? val : 0u);
                    break;
                }
                prefix += __reduce_add_sync(FULL_MASK, val);
                base -= 32;
            }
            if (lane == 0)
                Desc<uint32_t>::publish(desc, tile, DESC_PREFIX, prefix + total);
        }
        if (tile == ntiles - 1 && lane == 0)
            *count_out = prefix + total;
        if (lane < ENTRIES)
            s_excl[lane] = prefix + a - s_incl[lane];
    }
    __syncthreads();

    // ---- compact each warp row in shared memory, then store contiguously
    uint32_t *stage = s_stage[warp];
    #pragma unroll
    for (int j = 0; j < J; ++j) {
        const uint32_t row_count = __shfl_sync(FULL_MASK, incl[j], 31);
        if (row_count == 0)
            continue; // warp-uniform
        const uint32_t row_base = s_excl[j * COMPRESS_WARPS + warp];
        // item index of this vector's first byte (may wrap below zero for the
        // masked-out head bytes, which are never selected)
        const uint32_t item0 = (uint32_t) ((uint64_t) tile * TILE +
                                           (uint64_t) (j * COMPRESS_THREADS + tid) * 16 - mis);
        uint32_t o = incl[j] - cnt[j];
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t bits = nz[j][w];
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (bits & (0x80u << (8 * k))) {
                    stage[o] = item0 + w * 4 + k;
                    o++;
                }
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < row_count; i += 32)
            out[row_base + i] = stage[i];
        __syncwarp();
    }
}


// --------------------------------------------------- bit-packed two-pass path
//
// Large masks.  A single-pass compaction chained by look-back is limited by the
// rate at which tiles can be chained (measured ~90 tiles/us with scan_fast.cu's
// machinery, i.e. 1.5 TB/s of mask for 16 KiB tiles), and 4 bytes of staging per
// mask byte cap the tile size.  Packing the mask to one BIT per entry first makes
// a second pass cheap instead: 1 + 1/8 + 1/8 + 4 d bytes per entry against the
// algorithmic 1 + 4 d, with three simple, fully parallel kernels:
//   pack    16 mask bytes per thread -> 16 bits; 32-bit words of the bit mask and
//           one count per 8192-entry tile                          (HBM: n + n/8)
//   offsets exclusive scan of the tile counts by one CTA          (L2-resident)
//   expand  one bit-mask word per thread, ranks by popcount + block scan; a warp
//           walks its non-empty words and stores every word's index run with
//           one contiguous store instruction                  (HBM: n/8 + 4 count)
// Only bit 0 of a mask byte is looked at (entries are required to be 0 or 1,
// jit.h:2377-2379).

static constexpr int CP_THREADS = 256;
static constexpr uint32_t CP_TILE = 8192;                  // entries per expand tile
static constexpr uint32_t CP_WORDS = CP_TILE / 32;         // 256 words = one per thread

/// bit k of the result = bit 0 of byte k of w (k < 4)
B200_DEVICE uint32_t pack4(uint32_t w) {
    return (((w & 0x01010101u) * 0x00204081u) >> 21) & 0xfu;
}

B200_DEVICE uint32_t pack16(uint4 v) {
    return pack4(v.x) | (pack4(v.y) << 4) | (pack4(v.z) << 8) | (pack4(v.w) << 12);
}

/// Virtual layout: entry i of the mask is byte (i + mis) of the 16-byte aligned
/// array (in - mis); bits / tiles are indexed in that virtual space.  A CTA packs
/// two tiles (16384 virtual bytes): warp w, step s covers the 1024 bytes starting
/// at ((2 * blockIdx.x * 8 + s * 8 + w) * 1024).
__global__ void __launch_bounds__(CP_THREADS)
compress_pack_kernel(const uint8_t *__restrict__ in, uint64_t size, uint32_t mis,
                     uint32_t *__restrict__ bits, uint32_t *__restrict__ counts,
                     uint32_t ntiles) {
    __shared__ uint32_t s_cnt[2][CP_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *vin = in - mis;
    const uint64_t vend = size + mis; // valid virtual bytes: [mis, vend)

    uint4 v[2][2];
    #pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint64_t chunk = ((uint64_t) blockIdx.x * 2 + s) * 8 + warp; // 1024-byte chunk
        #pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint64_t vb = chunk * 1024 + (uint64_t) r * 512 + lane * 16;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (vb >= mis && vb + 16 <= vend) {
                x = ld_stream(vin + vb);
            } else if (vb < vend && vb + 16 > mis) {
                uint32_t w[4] = { 0, 0, 0, 0 };
                #pragma unroll
                for (int b = 0; b < 16; ++b)
                    if (vb + b >= mis && vb + b < vend)
                        w[b >> 2] |= (uint32_t) vin[vb + b] << (8 * (b & 3));
                x = make_uint4(w[0], w[1], w[2], w[3]);
            }
            v[s][r] = x;
        }
    }
    #pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint64_t chunk = ((uint64_t) blockIdx.x * 2 + s) * 8 + warp;
        const uint32_t b0 = pack16(v[s][0]), b1 = pack16(v[s][1]);
        // even lanes assemble the word of row 0, odd lanes the word of row 1
        const uint32_t got = __shfl_xor_sync(FULL_MASK, (lane & 1) ? b0 : b1, 1);
        const uint32_t word = (lane & 1) ? (got | (b1 << 16)) : (b0 | (got << 16));
        const uint64_t widx = chunk * 32 + (lane & 1) * 16 + (lane >> 1);
        if (widx * 32 < vend)
            bits[widx] = word;
        const uint32_t c = __reduce_add_sync(FULL_MASK, __popc(word));
        if (lane == 0)
            s_cnt[s][warp] = c;
    }
    __syncthreads();
    if (tid < 2) {
        uint32_t c = 0;
        #pragma unroll
        for (int w = 0; w < CP_THREADS / 32; ++w)
            c += s_cnt[tid][w];
        const uint32_t tile = blockIdx.x * 2 + tid;
        if (tile < ntiles)
            counts[tile] = c;
    }
}

/// In-place exclusive scan of the tile counts by one CTA (16 counts per thread
/// and step, loaded up front); total -> *count_out.
__global__ void __launch_bounds__(1024)
compress_offsets_kernel(uint32_t *__restrict__ counts, uint32_t ntiles,
                        uint32_t *__restrict__ count_out) {
    constexpr int PER = 16;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024 * PER) {
        const uint32_t i = base + tid * PER;
        uint32_t c[PER];
        #pragma unroll
        for (int k = 0; k < PER; ++k)
            c[k] = i + k < ntiles ? counts[i + k] : 0;
        uint32_t mine = 0;
        #pragma unroll
        for (int k = 0; k < PER; ++k)
            mine += c[k];
        uint32_t incl = mine;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= (uint32_t) d)
                incl += up;
        }
        if (lane == 31)
            s_warp[warp] = incl;
        __syncthreads();
        uint32_t wsum = s_warp[lane], wincl = wsum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t up = __shfl_up_sync(FULL_MASK, wincl, d);
            if (lane >= (uint32_t) d)
                wincl += up;
        }
        const uint32_t carry = s_carry;
        uint32_t run = carry + __shfl_sync(FULL_MASK, wincl - wsum, warp) + incl - mine;
        #pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (i + k < ntiles)
                counts[i + k] = run;
            run += c[k];
        }
        __syncthreads();
        if (tid == 1023)
            s_carry = run;
        __syncthreads();
    }
    if (tid == 0)
        *count_out = s_carry;
}

/// CP_EXP consecutive tiles of 8192 entries per CTA, one bit-mask word per thread
/// and tile (all loaded up front).  After the block-wide exclusive scan of the
/// popcounts a warp walks its non-empty words: the word and its output offset
/// are broadcast, lane j owns bit j and stores the index at offset + (number of
/// set bits below j).  Set bits of a word land on consecutive addresses, so every
/// store instruction writes one contiguous run (no shared-memory staging).
static constexpr uint32_t CP_EXP = 4;

__global__ void __launch_bounds__(CP_THREADS)
compress_expand_kernel(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ offsets,
                       uint32_t ntiles, uint32_t mis, uint32_t *__restrict__ out) {
    __shared__ uint32_t s_warp[CP_EXP][CP_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile0 = blockIdx.x * CP_EXP;
    const uint32_t below = (1u << lane) - 1u;

    uint32_t word[CP_EXP], excl[CP_EXP];
    #pragma unroll
    for (uint32_t t = 0; t < CP_EXP; ++t)
        word[t] = tile0 + t < ntiles ? __ldg(bits + (uint64_t) (tile0 + t) * CP_WORDS + tid) : 0u;
    #pragma unroll
    for (uint32_t t = 0; t < CP_EXP; ++t) {
        const uint32_t cnt = __popc(word[t]);
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= (uint32_t) d)
                incl += up;
        }
        if (lane == 31)
            s_warp[t][warp] = incl;
        excl[t] = incl - cnt;
    }
    __syncthreads();
    #pragma unroll
    for (uint32_t t = 0; t < CP_EXP; ++t) {
        if (tile0 + t >= ntiles)
            break;
        uint32_t before = __ldg(offsets + tile0 + t);
        #pragma unroll
        for (int w = 0; w < CP_THREADS / 32; ++w)
            before += (uint32_t) w < warp ? s_warp[t][w] : 0u;
        const uint32_t first = before + excl[t]; // output slot of this word's first set bit
        const uint32_t item0 = (tile0 + t) * CP_TILE + warp * 1024 - mis + lane; // bit 'lane' of word 0
        const uint32_t nz = __ballot_sync(FULL_MASK, word[t] != 0);
        #pragma unroll 1
        for (uint32_t g = 0; g < 32; g += 4) {
            if (((nz >> g) & 0xfu) == 0)
                continue; // warp-uniform
            #pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t w = __shfl_sync(FULL_MASK, word[t], g + u);
                const uint32_t o = __shfl_sync(FULL_MASK, first, g + u);
                if ((w >> lane) & 1u)
                    out[o + __popc(w & below)] = item0 + (g + u) * 32;
            }
        }
    }
}

// ------------------------------------------------------- single-pass stream
//
// Large masks, one launch: the mask is read exactly once (1 byte per entry), the
// indices are written exactly once -- the algorithmic n + 4 count bytes.
//
// What makes a single pass cheap here is that a packed tile is tiny: a thread
// reduces its J 16-byte vectors (J * 16 entries) to J 16-bit words kept in J / 2
// registers, so a CTA can hold THREE tiles at once and run a software pipeline
// around the one unavoidable latency, the look-back:
//
//   compute warp, iteration k   its part of the mask of tile k arrives in shared
//                 memory (bulk copy issued an iteration ago) -> pack to bits ->
//                 request tile k + 1 -> publish the warp total (mbarrier arrive, no
//                 wait) -> EXPAND TILE k - 2, whose first output slots were resolved
//                 in the meantime.  No CTA-wide barrier.
//   resolve warp  (one extra warp, on its own clock) waits until all 16 warp totals
//                 of tile k are in, publishes the tile's count as a 64-bit {status,
//                 count} descriptor, reads the counts of the whole WAVE of tiles
//                 (below) for the exclusive prefix and hands every compute warp its
//                 first output slot through a second mbarrier.
//
// Persistent CTAs, two per SM, all co-resident (cooperative launch): tile k of CTA b
// is tile k * gridDim + b, i.e. all CTAs work on the same wave of gridDim
// consecutive tiles.  There is no chained look-back: a CTA reads the counts of ALL
// tiles of its wave with loads that are in flight together -- one round trip when
// nobody is late -- and derives its exclusive prefix inside the wave and the wave
// total, which every CTA accumulates on its own into the base of the next wave.
// Why: a dependent global load costs ~2.5 us in this kernel (it queues behind
// 19 MB of bulk copies in flight; phase time line in tools/cs_trace.py).  A
// decoupled look-back with 32-wide windows needed ~20 such polls per tile, and a
// version that read the wave in three rounds of loads still 7.5 us per wave.
//
// Layout of a tile (virtual bytes, i.e. relative to the 16-byte aligned address
// in - mis): warp w owns the contiguous bytes [w * J * 512, (w + 1) * J * 512),
// vector j of lane l covers bytes j * 512 + l * 16 ... + 16 of that range, so a
// load instruction of a warp reads 512 contiguous bytes.
//
// Expansion of a row (512 entries = one 16-bit word per lane):
//   dense   (> CS_SPARSE set entries): every lane appends the indices of its set
//           bits to a per-warp staging row in shared memory (16 predicated
//           stores, no popc / shuffle per entry); the row then leaves with ONE
//           bulk copy shared -> global (cp.async.bulk) for its 16-byte aligned
//           middle plus at most 3 + 3 scalar stores for the ragged ends.  The
//           previous version (broadcast word + offset, lane i owns bit i) cost
//           27 warp instructions per 32 entries and was issue bound (ncu: 70 %
//           issue active at density 0.5).
//   sparse  every lane walks the set bits of its own word (a few iterations)
//           and stores straight to global memory.

#ifdef B200_CS_TRACE
// development aid: per-CTA phase time stamps (globaltimer, ns) of warp 0 and of the
// resolving warp; [cta][iteration][event]
static constexpr int CST_ITERS = 16, CST_EVENTS = 12;
__device__ unsigned long long cs_trace[320][CST_ITERS][CST_EVENTS];
B200_DEVICE unsigned long long cs_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define CS_STAMP(k, e) do { if (lane == 0 && (k) < CST_ITERS) cs_trace[blockIdx.x][k][e] = cs_now(); } while (0)
#else
#define CS_STAMP(k, e) do { } while (0)
#endif

static constexpr int CS_THREADS = 512;             // compute threads (+ 32: resolve warp)
static constexpr int CS_WARPS = CS_THREADS / 32;
static constexpr int CS_WAVE_MAX = 320;            // CTAs per wave (10 descriptors per lane)
static constexpr uint32_t CS_SPARSE = 48;          // set entries per 512-entry row
static constexpr uint32_t CS_NONE = 0xffffffffu;
static constexpr uint32_t CS_STAGE = 512 + 4;      // staging words per warp (row + alignment slack)
static constexpr int CS_RING = 4;                  // look-back hand-over ring (tiles)

template <int J>
__global__ void __launch_bounds__(CS_THREADS + 32, 2)
compress_stream_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t size,
                       uint32_t mis, uint32_t ntiles, uint64_t *desc, uint32_t *count_out) {
    static_assert(J % 2 == 0 && J <= 8, "J 16-bit words are kept in J / 2 registers");
    constexpr uint32_t WARP_BYTES = J * 512;
    constexpr uint32_t TILE = CS_WARPS * WARP_BYTES;

    __shared__ uint32_t s_wtot[CS_RING][CS_WARPS];  // set entries per warp of tile k in [k % RING]
    __shared__ uint32_t s_first[CS_RING][CS_WARPS]; // first output slot per warp of tile k
    __shared__ __align__(8) uint64_t s_full[CS_WARPS];   // the warp's part of a tile has landed
    __shared__ __align__(8) uint64_t s_tot_bar[CS_RING]; // totals of tile k are complete (16 arrivals)
    __shared__ __align__(8) uint64_t s_pfx_bar[CS_RING]; // first slots of tile k are resolved
    extern __shared__ __align__(128) uint8_t cs_smem[];  // CS_WARPS * (WARP_BYTES + CS_STAGE * 4)

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *vin = in - mis;
    const uint64_t vend = size + mis; // valid virtual bytes: [mis, vend)

    if (tid == 0) {
        #pragma unroll
        for (int w = 0; w < CS_WARPS; ++w)
            mbar_init(&s_full[w], 1);
        #pragma unroll
        for (int r = 0; r < CS_RING; ++r) {
            mbar_init(&s_tot_bar[r], CS_WARPS);
            mbar_init(&s_pfx_bar[r], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto tile_of = [&](uint32_t k) -> uint32_t {
        const uint64_t t = (uint64_t) k * gridDim.x + blockIdx.x;
        return t < ntiles ? (uint32_t) t : CS_NONE;
    };

    // ---- resolve warp
    if (warp == CS_WARPS) {
        uint32_t base = 0; // total of all earlier waves
        for (uint32_t k = 0;; ++k) {
            const uint32_t tile = tile_of(k);
            if (tile == CS_NONE)
                break;
            const uint32_t slot = k % CS_RING;
            CS_STAMP(k, 4);
            mbar_wait(&s_tot_bar[slot], (k / CS_RING) & 1);
            CS_STAMP(k, 5);
            const uint32_t mine = lane < CS_WARPS ? s_wtot[slot][lane] : 0u;
            uint32_t incl = mine;
            #pragma unroll
            for (int d = 1; d < CS_WARPS; d <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= (uint32_t) d)
                    incl += up;
            }
            const uint32_t total = __shfl_sync(FULL_MASK, incl, CS_WARPS - 1);
            if (lane == 0)
                Desc<uint32_t>::publish(desc, tile, DESC_AGGREGATE, total);

            // counts of the whole wave: every load is issued before the first one is
            // looked at; entries that are not there yet are polled again
            const uint32_t wave0 = tile - blockIdx.x;               // first tile of the wave
            const uint32_t wave_n = min(gridDim.x, ntiles - wave0); // tiles in this wave
            constexpr int NQ = CS_WAVE_MAX / 32;
            uint32_t val[NQ], st[NQ];
            #pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const uint32_t j = 32 * q + lane;
                val[q] = 0;
                st[q] = DESC_AGGREGATE;
                if (j < wave_n)
                    st[q] = Desc<uint32_t>::observe(desc, wave0 + j, val[q]);
            }
            uint32_t below = 0, all = 0;
#ifdef B200_CS_TRACE
            uint32_t repolls = 0;
            {
                uint32_t any = 0;
                #pragma unroll
                for (int q = 0; q < NQ; ++q)
                    any |= st[q];
                if (__any_sync(FULL_MASK, any == 12345u)) // forces the loads to complete
                    repolls = 1000;
                CS_STAMP(k, 3);
            }
#endif
            #pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const uint32_t j = 32 * q + lane;
                while (__any_sync(FULL_MASK, st[q] == DESC_INVALID)) {
#ifdef B200_CS_TRACE
                    repolls++;
#endif
                    if (st[q] == DESC_INVALID)
                        st[q] = Desc<uint32_t>::observe(desc, wave0 + j, val[q]);
                }
                all += val[q];
                below += j < blockIdx.x ? val[q] : 0u;
            }
            all = __reduce_add_sync(FULL_MASK, all);
            below = __reduce_add_sync(FULL_MASK, below);
            CS_STAMP(k, 6);
#ifdef B200_CS_TRACE
            if (lane == 0 && k < CST_ITERS)
                cs_trace[blockIdx.x][k][11] = repolls;
#endif
            const uint32_t prefix = base + below;
            base += all;
            if (lane < CS_WARPS)
                s_first[slot][lane] = prefix + incl - mine;
            if (lane == 0 && tile == ntiles - 1)
                *count_out = prefix + total;
            __syncwarp();
            CS_STAMP(k, 7);
            if (lane == 0)
                mbar_arrive(&s_pfx_bar[slot]);
        }
        return;
    }

    // ---- compute warps
    // The mask of a tile reaches shared memory through one 1-D bulk copy per warp
    // (the warp's J * 512 bytes are contiguous), issued by lane 0 an iteration
    // ahead and tracked by the warp's own mbarrier.  Tiles that are not entirely
    // inside the array (a misaligned head, the tail) are read with guarded loads.
    uint4 *wbuf = (uint4 *) cs_smem + (size_t) warp * (WARP_BYTES / 16);
    uint32_t *stage = (uint32_t *) (cs_smem + (size_t) CS_WARPS * WARP_BYTES) + (size_t) warp * CS_STAGE;
    auto tile_is_full = [&](uint32_t tile) {
        return (uint64_t) tile * TILE >= mis && (uint64_t) (tile + 1) * TILE <= vend;
    };
    auto request_tile = [&](uint32_t tile) { // lane 0 only
        if (tile_is_full(tile)) {
            mbar_arrive_expect_tx(&s_full[warp], WARP_BYTES);
            bulk_g2s(wbuf, vin + (uint64_t) tile * TILE + warp * WARP_BYTES, WARP_BYTES, &s_full[warp]);
        } else {
            mbar_arrive(&s_full[warp]);
        }
    };
    auto fetch_vector = [&](uint32_t tile, bool full, int j) -> uint4 {
        if (full)
            return wbuf[j * 32 + lane];
        const uint64_t vb = (uint64_t) tile * TILE + warp * WARP_BYTES + j * 512 + lane * 16;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (vb >= mis && vb + 16 <= vend) {
            x = ld_stream(vin + vb);
        } else if (vb < vend && vb + 16 > mis) {
            uint32_t w[4] = { 0, 0, 0, 0 };
            #pragma unroll
            for (int b = 0; b < 16; ++b)
                if (vb + b >= mis && vb + b < vend)
                    w[b >> 2] |= (uint32_t) vin[vb + b] << (8 * (b & 3));
            x = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return x;
    };

    // expansion of one tile held as packed bits (hp[q] = word of row 2q | word of
    // row 2q + 1 << 16); 'first' = output slot of the warp's first set entry
    bool store_pending = false; // a bulk store may still be reading the staging row
    auto expand = [&](uint32_t tile, const uint32_t (&hp)[J / 2], uint32_t first) {
        // inclusive scans over the lanes of all J row counts at once: three
        // 10-bit fields per register (a row holds at most 512 set entries)
        constexpr int NP = (J + 2) / 3;
        uint32_t c[J], P[NP];
        #pragma unroll
        for (int j = 0; j < J; ++j)
            c[j] = __popc((hp[j / 2] >> (16 * (j & 1))) & 0xffffu);
        #pragma unroll
        for (int i = 0; i < NP; ++i) {
            P[i] = 0;
            #pragma unroll
            for (int f = 0; f < 3; ++f)
                if (i * 3 + f < J)
                    P[i] |= c[i * 3 + f] << (10 * f);
        }
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            #pragma unroll
            for (int i = 0; i < NP; ++i) {
                const uint32_t up = __shfl_up_sync(FULL_MASK, P[i], d);
                if (lane >= (uint32_t) d)
                    P[i] += up;
            }
        }
        uint32_t T[NP];
        #pragma unroll
        for (int i = 0; i < NP; ++i)
            T[i] = __shfl_sync(FULL_MASK, P[i], 31);

        // entry index of bit 0 of this lane's word of row 0 (may wrap below zero
        // for the masked-out head bytes, which are never set)
        const uint32_t item0 = (uint32_t) ((uint64_t) tile * TILE + warp * WARP_BYTES + lane * 16 - mis);
        uint32_t row_first = first;
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t row_total = (T[j / 3] >> (10 * (j % 3))) & 0x3ffu;
            if (row_total == 0)
                continue; // warp-uniform
            const uint32_t incl = (P[j / 3] >> (10 * (j % 3))) & 0x3ffu;
            uint32_t word = (hp[j / 2] >> (16 * (j & 1))) & 0xffffu;
            const uint32_t item = item0 + j * 512;
            if (row_total <= CS_SPARSE) {
                uint32_t o = row_first + incl - c[j]; // slot of this lane's first set entry
                while (word) {
                    const uint32_t b = __ffs(word) - 1;
                    word &= word - 1;
                    out[o++] = item + b;
                }
            } else {
                // stage[a + i] <-> out[row_first + i]: 'a' makes 16-byte aligned
                // staging words coincide with 16-byte aligned global addresses
                uint32_t *dst = out + row_first;
                const uint32_t a = (uint32_t) (((uintptr_t) dst >> 2) & 3u);
                if (store_pending) {
                    if (lane == 0)
                        bulk_wait_read<0>();
                    __syncwarp();
                }
                uint32_t *sp = stage + a + (incl - c[j]);
                #pragma unroll
                for (int b = 0; b < 16; ++b) {
                    if (word & (1u << b))
                        *sp++ = item + b;
                }
                fence_proxy_async();
                __syncwarp();
                const uint32_t end = a + row_total;           // staging words [a, end) are valid
                const uint32_t lo = (a + 3u) & ~3u, hi = end & ~3u;
                if (hi > lo) {
                    if (lane == 0) {
                        bulk_s2g(dst - a + lo, stage + lo, (hi - lo) * 4);
                        bulk_commit();
                    }
                    store_pending = true;
                    // ragged ends: [a, lo) and [hi, end)
                    if (lane < lo - a)
                        dst[lane] = stage[a + lane];
                    else if (lane >= 4 && lane - 4 < end - hi)
                        dst[hi - a + lane - 4] = stage[hi + lane - 4];
                } else {
                    // fewer than one aligned vector (cannot happen for dense rows,
                    // kept for safety)
                    for (uint32_t i = lane; i < row_total; i += 32)
                        dst[i] = stage[a + i];
                    __syncwarp();
                }
            }
            row_first += row_total;
        }
    };

    // tiles k - 1 and k - 2 wait for their prefix as packed bits in registers
    uint32_t hp1[J / 2], hp2[J / 2], tile1 = CS_NONE, tile2 = CS_NONE;
    #pragma unroll
    for (int q = 0; q < J / 2; ++q)
        hp1[q] = hp2[q] = 0;
    if (lane == 0 && tile_of(0) != CS_NONE)
        request_tile(tile_of(0));
    for (uint32_t k = 0;; ++k) {
        const uint32_t tile = tile_of(k);
        uint32_t hp[J / 2];
        #pragma unroll
        for (int q = 0; q < J / 2; ++q)
            hp[q] = 0;
        if (tile != CS_NONE) {
            const bool full = tile_is_full(tile);
            if (warp == 0) CS_STAMP(k, 0);
            mbar_wait(&s_full[warp], k & 1);
            if (warp == 0) CS_STAMP(k, 1);
            uint32_t cnt = 0;
            #pragma unroll
            for (int q = 0; q < J / 2; ++q) {
                hp[q] = pack16(fetch_vector(tile, full, 2 * q)) |
                        (pack16(fetch_vector(tile, full, 2 * q + 1)) << 16);
                cnt += __popc(hp[q]);
            }
            __syncwarp(); // every lane has read the buffer: refill it
            const uint32_t next = tile_of(k + 1);
            if (lane == 0 && next != CS_NONE)
                request_tile(next);
            const uint32_t wtotal = __reduce_add_sync(FULL_MASK, cnt);
            if (lane == 0) {
                // slot k % RING was last read for tile k - RING, which every warp
                // has expanded (it waited for that tile's prefix two iterations ago)
                s_wtot[k % CS_RING][warp] = wtotal;
                mbar_arrive(&s_tot_bar[k % CS_RING]);
            }
            if (warp == 0) CS_STAMP(k, 2);
        }
        if (tile2 != CS_NONE) {
            const uint32_t slot = (k - 2) % CS_RING;
            if (warp == 0) CS_STAMP(k, 8);
            mbar_wait(&s_pfx_bar[slot], ((k - 2) / CS_RING) & 1);
            if (warp == 0) CS_STAMP(k, 9);
            expand(tile2, hp2, s_first[slot][warp]);
            if (warp == 0) CS_STAMP(k, 10);
        }
        if (tile == CS_NONE && tile1 == CS_NONE)
            break;
        #pragma unroll
        for (int q = 0; q < J / 2; ++q) {
            hp2[q] = hp1[q];
            hp1[q] = hp[q];
        }
        tile2 = tile1;
        tile1 = tile;
    }
    // the staging row must outlive the last bulk store that reads it
    if (store_pending && lane == 0)
        bulk_wait_read<0>();
}

static int compress_two_pass(cudaStream_t stream, const uint8_t *in, uint64_t size, uint32_t *out,
                             uint32_t *count_dev);

static int compress_stream(cudaStream_t stream, const uint8_t *in, uint64_t size, uint32_t *out,
                           uint32_t *count_dev) {
    constexpr int J = 8;
    constexpr uint32_t TILE = CS_WARPS * J * 512;
    constexpr size_t smem = TILE + (size_t) CS_WARPS * CS_STAGE * 4;
    auto kernel = compress_stream_kernel<J>;
    uint32_t mis = (uint32_t) ((uintptr_t) in & 15);
    uint32_t ntiles = (uint32_t) ceil_div(size + mis, TILE);

    // co-residency of all CTAs is what guarantees forward progress of the look-back:
    // cooperative launch, grid <= what fits; without support for it, two passes
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = dev < 64 ? occ_cache[dev].load(std::memory_order_relaxed) : 0;
    if (occ == 0) {
        int coop = 0;
        B200_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CS_THREADS + 32, smem));
        occ = coop && occ >= 1 ? occ : -1;
        if (dev < 64)
            occ_cache[dev].store(occ, std::memory_order_relaxed);
    }
    if (occ < 0)
        return compress_two_pass(stream, in, size, out, count_dev);

    const size_t desc_bytes = (size_t) ntiles * sizeof(uint64_t);
    uint64_t *desc = (uint64_t *) temp_alloc(desc_bytes, stream);
    if (!desc)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory");
    cudaError_t err = cudaMemsetAsync(desc, 0, desc_bytes, stream);
    if (err != cudaSuccess) {
        temp_free(desc, stream);
        return cuda_fail(err, "cudaMemsetAsync");
    }
    const uint32_t grid = (uint32_t) std::min<uint64_t>(
        std::min<uint64_t>(ntiles, (uint64_t) sm_count() * occ), CS_WAVE_MAX);
    void *args[] = { (void *) &in, (void *) &out, (void *) &size, (void *) &mis, (void *) &ntiles,
                     (void *) &desc, (void *) &count_dev };
    err = cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(CS_THREADS + 32), args, smem, stream);
    temp_free(desc, stream);
    if (err != cudaSuccess)
        return cuda_fail(err, "cudaLaunchCooperativeKernel(compress_stream_kernel)");
    count_launch();
    return B200_OK;
}

/// 1 (default): single-pass stream kernel, 2: bit-packed two-pass path
/// (development switch, kept for A/B measurements)
static int compress_path() {
    static int path = -1;
    if (path < 0) {
        const char *s = getenv("B200_COMPRESS_PATH");
        path = s ? atoi(s) : 1;
    }
    return path;
}

static int compress_two_pass(cudaStream_t stream, const uint8_t *in, uint64_t size, uint32_t *out,
                             uint32_t *count_dev) {
    const uint32_t mis = (uint32_t) ((uintptr_t) in & 15);
    const uint64_t vsize = size + mis;
    const uint32_t ntiles = (uint32_t) ceil_div(vsize, CP_TILE);
    const uint64_t nwords = (uint64_t) ntiles * CP_WORDS;
    const size_t bytes = (size_t) (nwords + ntiles) * sizeof(uint32_t);
    uint32_t *bits = (uint32_t *) temp_alloc(bytes, stream);
    if (!bits)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory (%zu bytes)", bytes);
    uint32_t *counts = bits + nwords;
    // words past the end of the mask are never written by the pack kernel
    const uint64_t last_word = ceil_div(vsize, 32);
    if (nwords > last_word) {
        cudaError_t err = cudaMemsetAsync(bits + last_word, 0, (nwords - last_word) * 4, stream);
        if (err != cudaSuccess) {
            temp_free(bits, stream);
            return cuda_fail(err, "cudaMemsetAsync");
        }
    }
    compress_pack_kernel<<<(ntiles + 1) / 2, CP_THREADS, 0, stream>>>(in, size, mis, bits, counts, ntiles);
    compress_offsets_kernel<<<1, 1024, 0, stream>>>(counts, ntiles, count_dev);
    compress_expand_kernel<<<(uint32_t) ceil_div(ntiles, CP_EXP), CP_THREADS, 0, stream>>>(bits, counts, ntiles, mis, out);
    count_launch(2);
    temp_free(bits, stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

#ifdef B200_CS_TRACE
__attribute__((visibility("default"))) int b200_debug_cs_trace(void *dst, size_t bytes) {
    return (int) cudaMemcpyFromSymbol(dst, cs_trace, std::min(bytes, sizeof(cs_trace)));
}
#endif

int b200_compress_async(void *stream_, const uint8_t *in, uint64_t size, uint32_t *out,
                        uint32_t *count_dev) {
    int rc = ensure_init();
    if (rc)
        return rc;
    cudaStream_t stream = resolve_stream(stream_);
    if (size == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), stream));
        return B200_OK;
    }
    if (size > 0xffffffffull)
        return fail(B200_ERR_INVALID, "jit_compress(): array too large (indices are 32 bit)!");
    // small masks: one launch of the single-pass kernel; large ones: two passes
    // over a bit-packed copy of the mask
    if (size > 32768)
        return compress_path() == 2 ? compress_two_pass(stream, in, size, out, count_dev)
                                    : compress_stream(stream, in, size, out, count_dev);

    constexpr int J = 4;
    constexpr uint32_t TILE = COMPRESS_THREADS * J * 16;
    uint32_t mis = (uint32_t) ((uintptr_t) in & 15);
    uint32_t ntiles = (uint32_t) ceil_div(size + mis, TILE);
    size_t desc_bytes = (size_t) ntiles * sizeof(uint64_t);
    void *scratch = temp_alloc(desc_bytes + 16, stream);
    if (!scratch)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory");
    B200_CUDA_CHECK(cudaMemsetAsync(scratch, 0, desc_bytes + 16, stream));
    compress_kernel<J><<<ntiles, COMPRESS_THREADS, 0, stream>>>(
        in, out, size, ntiles, (uint64_t *) scratch,
        (uint32_t *) ((uint8_t *) scratch + desc_bytes), count_dev);
    temp_free(scratch, stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int b200_compress(void *stream_, const uint8_t *in, uint64_t size, uint32_t *out,
                  uint32_t *count) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (size == 0) { // src/cuda_ts.cpp:685-686
        *count = 0;
        return B200_OK;
    }
    cudaStream_t stream = resolve_stream(stream_);
    // the reference reads the count from pinned memory after a full stream
    // synchronisation (src/cuda_ts.cpp:759-762); here the kernel writes it there
    uint32_t *count_pinned = pinned_scalar();
    if (!count_pinned)
        return fail(B200_ERR_CUDA, "jit_compress(): could not allocate pinned memory");
    rc = b200_compress_async(stream, in, size, out, count_pinned);
    if (rc)
        return rc;
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    *count = *count_pinned;
    return B200_OK;
}

} // extern "C"
