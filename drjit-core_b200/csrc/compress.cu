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
// registers, so a CTA can hold TWO tiles at once and run a software pipeline
// around the one unavoidable latency, the look-back:
//
//   iteration k   compute warps: mask of tile k arrives in registers (loads were
//                 issued an iteration ago) -> pack to bits -> issue the loads of
//                 tile k + 1 -> publish the warp totals -> barrier -> EXPAND
//                 TILE k - 1, whose prefix was resolved in the meantime;
//                 look-back warp: after the same barrier publishes the aggregate
//                 of tile k, resolves its exclusive prefix over 64-bit
//                 {status, count} descriptors and leaves it for iteration k + 1.
//
// Persistent CTAs (two per SM), tiles handed out by an atomic ticket two
// iterations ahead (forward progress of the look-back + load balance).
//
// Layout of a tile (virtual bytes, i.e. relative to the 16-byte aligned address
// in - mis): warp w owns the contiguous bytes [w * J * 512, (w + 1) * J * 512),
// vector j of lane l covers bytes j * 512 + l * 16 ... + 16 of that range, so a
// load instruction of a warp reads 512 contiguous bytes.
//
// Expansion of a row (512 entries = one 16-bit word per lane):
//   dense   (> CS_SPARSE set entries): the row is walked two lanes' words at a
//           time; the words and their output offsets are broadcast, lane i owns
//           bit i & 15 of word i >> 4 and stores its index at offset + (number of
//           set bits below it): every store instruction writes one or two
//           contiguous runs.
//   sparse  every lane walks the set bits of its own word (a few iterations).

static constexpr int CS_THREADS = 512;             // compute threads (+ 32: look-back warp)
static constexpr int CS_WARPS = CS_THREADS / 32;
static constexpr uint32_t CS_SPARSE = 48;          // set entries per 512-entry row
static constexpr uint32_t CS_NONE = 0xffffffffu;

template <int J>
__global__ void __launch_bounds__(CS_THREADS + 32, 2)
compress_stream_kernel(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t size,
                       uint32_t mis, uint32_t ntiles, uint64_t *desc, uint32_t *ticket,
                       uint32_t *count_out) {
    static_assert(J % 2 == 0 && J <= 8, "J 16-bit words are kept in J / 2 registers");
    constexpr uint32_t WARP_BYTES = J * 512;
    constexpr uint32_t TILE = CS_WARPS * WARP_BYTES;

    __shared__ uint32_t s_tile[4];            // tile of iteration k in s_tile[k % 4]
    __shared__ uint32_t s_wtot[2][CS_WARPS];  // set entries per warp of tile k in [k & 1]
    __shared__ uint32_t s_prefix[2];          // exclusive prefix of tile k in [k & 1]
    __shared__ __align__(8) uint64_t s_full[CS_WARPS]; // the warp's part of a tile has landed
    extern __shared__ __align__(128) uint8_t cs_smem[]; // CS_WARPS * WARP_BYTES

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *vin = in - mis;
    const uint64_t vend = size + mis; // valid virtual bytes: [mis, vend)

    if (tid == 0) {
        #pragma unroll
        for (int w = 0; w < CS_WARPS; ++w)
            mbar_init(&s_full[w], 1);
        mbar_fence_init();
        const uint32_t t0 = atomicAdd(ticket, 1u), t1 = atomicAdd(ticket, 1u);
        s_tile[0] = t0 < ntiles ? t0 : CS_NONE;
        s_tile[1] = t1 < ntiles ? t1 : CS_NONE;
    }
    __syncthreads();

    // ---- look-back warp
    if (warp == CS_WARPS) {
        for (uint32_t k = 0;; ++k) {
            __syncthreads(); // warp totals of tile k are in s_wtot[k & 1]
            const uint32_t tile = s_tile[k % 4];
            if (tile == CS_NONE)
                break;
            const uint32_t total = __reduce_add_sync(FULL_MASK, lane < CS_WARPS ? s_wtot[k & 1][lane] : 0u);
            uint32_t prefix = 0;
            if (tile == 0) {
                if (lane == 0)
                    Desc<uint32_t>::publish(desc, 0, DESC_PREFIX, total);
            } else {
                if (lane == 0)
                    Desc<uint32_t>::publish(desc, tile, DESC_AGGREGATE, total);
                int64_t win = (int64_t) tile - 1;
                while (true) {
                    // window of the 32 preceding tiles, lane 0 nearest; only lanes
                    // whose entry is still INVALID poll again, entries beyond the
                    // nearest PREFIX are not waited for
                    const int64_t idx = win - lane;
                    uint32_t val = 0, st = DESC_PREFIX, pre;
                    if (idx >= 0)
                        st = Desc<uint32_t>::observe(desc, (uint32_t) idx, val);
                    while (true) {
                        pre = __ballot_sync(FULL_MASK, st == DESC_PREFIX);
                        uint32_t inv = __ballot_sync(FULL_MASK, st == DESC_INVALID);
                        if (pre)
                            inv &= (1u << (__ffs(pre) - 1)) - 1u;
                        if (!inv)
                            break;
                        __nanosleep(64);
                        if (st == DESC_INVALID)
                            st = Desc<uint32_t>::observe(desc, (uint32_t) idx, val);
                    }
                    if (pre) {
                        const uint32_t stop = __ffs(pre) - 1;
                        prefix += __reduce_add_sync(FULL_MASK, lane <= stop ? val : 0u);
                        break;
                    }
                    prefix += __reduce_add_sync(FULL_MASK, val);
                    win -= 32;
                }
                if (lane == 0)
                    Desc<uint32_t>::publish(desc, tile, DESC_PREFIX, prefix + total);
            }
            if (lane == 0) {
                s_prefix[k & 1] = prefix;
                if (tile == ntiles - 1)
                    *count_out = prefix + total;
            }
        }
        return;
    }

    // ---- compute warps
    // The mask of a tile reaches shared memory through one 1-D bulk copy per warp
    // (the warp's J * 512 bytes are contiguous), issued by lane 0 an iteration
    // ahead and tracked by the warp's own mbarrier.  Tiles that are not entirely
    // inside the array (a misaligned head, the tail) are read with guarded loads.
    uint4 *wbuf = (uint4 *) cs_smem + (size_t) warp * (WARP_BYTES / 16);
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
        const uint32_t sub = lane & 15, below = (1u << sub) - 1u;
        uint32_t row_first = first;
        #pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t row_total = (T[j / 3] >> (10 * (j % 3))) & 0x3ffu;
            if (row_total == 0)
                continue; // warp-uniform
            const uint32_t incl = (P[j / 3] >> (10 * (j % 3))) & 0x3ffu;
            uint32_t word = (hp[j / 2] >> (16 * (j & 1))) & 0xffffu;
            uint32_t o = row_first + incl - c[j]; // slot of this lane's first set entry
            if (row_total <= CS_SPARSE) {
                uint32_t item = item0 + j * 512;
                while (word) {
                    const uint32_t b = __ffs(word) - 1;
                    word &= word - 1;
                    out[o++] = item + b;
                }
            } else {
                const uint32_t nz = __ballot_sync(FULL_MASK, word != 0);
                // item of bit 'sub' of the word of lane (lane >> 4)
                const uint32_t item = item0 - lane * 16 + j * 512 + (lane >> 4) * 16 + sub;
                #pragma unroll 4
                for (uint32_t s = 0; s < 16; ++s) {
                    if (((nz >> (2 * s)) & 3u) == 0)
                        continue; // warp-uniform
                    const uint32_t src = 2 * s + (lane >> 4);
                    const uint32_t w = __shfl_sync(FULL_MASK, word, src);
                    const uint32_t wo = __shfl_sync(FULL_MASK, o, src);
                    if ((w >> sub) & 1u)
                        out[wo + __popc(w & below)] = item + s * 32;
                }
            }
            row_first += row_total;
        }
    };

    uint32_t hp_prev[J / 2], first_prev = 0, tile_prev = CS_NONE;
    if (lane == 0 && s_tile[0] != CS_NONE)
        request_tile(s_tile[0]);
    for (uint32_t k = 0;; ++k) {
        const uint32_t tile = s_tile[k % 4];
        // ticket of iteration k + 2: requested now, published before the barrier
        uint32_t t2 = 0;
        if (tid == 0 && tile != CS_NONE)
            t2 = atomicAdd(ticket, 1u);

        uint32_t hp[J / 2];
        if (tile != CS_NONE) {
            const bool full = tile_is_full(tile);
            mbar_wait(&s_full[warp], k & 1);
            uint32_t cnt = 0;
            #pragma unroll
            for (int q = 0; q < J / 2; ++q) {
                hp[q] = pack16(fetch_vector(tile, full, 2 * q)) |
                        (pack16(fetch_vector(tile, full, 2 * q + 1)) << 16);
                cnt += __popc(hp[q]);
            }
            __syncwarp(); // every lane has read the buffer: refill it
            const uint32_t next = s_tile[(k + 1) % 4];
            if (lane == 0 && next != CS_NONE)
                request_tile(next);
            const uint32_t wtotal = __reduce_add_sync(FULL_MASK, cnt);
            if (lane == 0)
                s_wtot[k & 1][warp] = wtotal;
            if (tid == 0)
                s_tile[(k + 2) % 4] = t2 < ntiles ? t2 : CS_NONE;
        }
        __syncthreads();

        if (tile_prev != CS_NONE)
            expand(tile_prev, hp_prev, first_prev + s_prefix[(k - 1) & 1]);
        if (tile == CS_NONE)
            break;

        uint32_t wexcl = 0;
        #pragma unroll
        for (int w = 0; w < CS_WARPS; ++w)
            wexcl += (uint32_t) w < warp ? s_wtot[k & 1][w] : 0u;
        #pragma unroll
        for (int q = 0; q < J / 2; ++q)
            hp_prev[q] = hp[q];
        first_prev = wexcl;
        tile_prev = tile;
    }
}

static int compress_stream(cudaStream_t stream, const uint8_t *in, uint64_t size, uint32_t *out,
                           uint32_t *count_dev) {
    constexpr int J = 8;
    constexpr uint32_t TILE = CS_WARPS * J * 512;
    auto kernel = compress_stream_kernel<J>;
    const uint32_t mis = (uint32_t) ((uintptr_t) in & 15);
    const uint32_t ntiles = (uint32_t) ceil_div(size + mis, TILE);

    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = dev < 64 ? occ_cache[dev].load(std::memory_order_relaxed) : 0;
    if (occ == 0) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TILE));
        B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CS_THREADS + 32, TILE));
        occ = occ < 1 ? 1 : occ;
        if (dev < 64)
            occ_cache[dev].store(occ, std::memory_order_relaxed);
    }

    const size_t desc_bytes = (size_t) ntiles * sizeof(uint64_t);
    void *scratch = temp_alloc(desc_bytes + 16, stream);
    if (!scratch)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory");
    cudaError_t err = cudaMemsetAsync(scratch, 0, desc_bytes + 16, stream);
    if (err != cudaSuccess) {
        temp_free(scratch, stream);
        return cuda_fail(err, "cudaMemsetAsync");
    }
    const uint32_t grid = (uint32_t) std::min<uint64_t>(ntiles, (uint64_t) sm_count() * occ);
    kernel<<<grid, CS_THREADS + 32, TILE, stream>>>(in, out, size, mis, ntiles, (uint64_t *) scratch,
                                                 (uint32_t *) ((uint8_t *) scratch + desc_bytes),
                                                 count_dev);
    temp_free(scratch, stream);
    B200_LAUNCH_CHECK();
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
    uint32_t *count_dev = (uint32_t *) temp_alloc(sizeof(uint32_t), stream);
    if (!count_dev)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory");
    rc = b200_compress_async(stream, in, size, out, count_dev);
    if (rc) {
        temp_free(count_dev, stream);
        return rc;
    }
    // the reference reads the count from pinned memory after a full stream
    // synchronisation (src/cuda_ts.cpp:759-762)
    cudaError_t err = cudaMemcpyAsync(count, count_dev, sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, stream);
    temp_free(count_dev, stream);
    if (err != cudaSuccess)
        return cuda_fail(err, "cudaMemcpyAsync");
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    return B200_OK;
}

} // extern "C"
