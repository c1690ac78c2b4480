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
// Larger masks take the bit-packed tile path further down.
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


// ------------------------------------------------ large masks: bit-packed tiles
//
// Two launches, no inter-CTA dependency inside either of them:
//   pack     the mask is read once (1 byte per entry, 128 bytes in flight per thread)
//            and reduced to one BIT per entry; a CTA also leaves the number of set
//            entries of its 32768-entry tile and adds it to the counter of its
//            group of 128 tiles                                   (HBM: n + n / 8)
//   expand   a CTA re-reads the 4 KiB of bits of its tile (L2 resident: the bit
//            array of a 2^28 mask is 32 MiB), sums the counters of the groups and
//            of the tiles of its own group in front of it (<= 255 + n / 2^21 words,
//            one or two loads per thread -- a separate scan kernel over the tile
//            counts cost 18 us) and writes the indices        (n / 8 + 4 count)
// i.e. 1.25 n + 4 count bytes against the algorithmic n + 4 count.
//
// A single pass (n + 4 count) needs every tile's prefix before its indices can be
// written.  Two such kernels were built and measured here -- decoupled look-back
// over {status, count} descriptors, and a cooperative "wave-synchronous" variant
// in which all resident CTAs exchange the counts of one wave of tiles -- with
// persistent CTAs, bulk-copy (TMA) loads and two tiles of slack.  Both ran at
// 1.8 - 2.1 TB/s of mask: a dependent global load issued behind ~19 MB of bulk
// copies in flight returns after 0.5 - 10 us (phase time line:
// profiles/r1_compress_stream_phase_trace.txt), so one hand-shake per 19 MB wave
// costs three times what streaming the wave costs.  The two-pass form has no
// hand-shake to wait for.
//
// Bit layout = the thread layout of both kernels: warp w of a tile owns entries
// [w * 4096, (w + 1) * 4096) as eight rows of 512; lane l holds one 16-bit word per
// row (entries l * 16 .. + 15 of the row), i.e. one 16-byte {rows 0|1, .., rows 6|7}
// vector per thread, stored / loaded fully coalesced.  (Four rows per warp: the
// fixed cost per warp -- lane scan, tile offset -- made the sparse case issue bound.)  Only bit 0 of a mask byte is
// looked at (entries are required to be 0 or 1, jit.h:2377-2379).
//
// Expansion of a warp's 4096 entries (eight 16-bit words per lane):
//   dense   (> 8 * CT_SPARSE set entries): every lane appends the indices of its set
//           bits to the warp's staging area in shared memory (16 predicated stores
//           per word, no popc / shuffle per entry), two rows at a time; the up to
//           4 KiB then leave with ONE bulk copy shared -> global (cp.async.bulk)
//           for the 16-byte aligned middle plus at most 3 + 3 scalar stores for the
//           ragged ends.  (Staging all four rows at once needs fewer instructions
//           but 66 KB per CTA: 3 CTAs per SM, measured slower.)  (An
//           earlier version -- broadcast word + offset, lane i owns bit i -- cost
//           27 warp instructions per 32 entries and was issue bound.)
//   sparse  every lane walks the set bits of its own words (a few iterations) and
//           stores straight to global memory.

static constexpr int CT_THREADS = 256;
static constexpr int CT_WARPS = CT_THREADS / 32;
static constexpr int CT_ROWS = 8;                                   // rows of 512 entries per warp
static constexpr uint32_t CT_WARP_ENTRIES = CT_ROWS * 512;          // 4096
static constexpr uint32_t CT_TILE = CT_WARPS * CT_WARP_ENTRIES;     // 32768 entries per CTA
struct __align__(16) CtBits { uint32_t w[CT_ROWS / 2]; };            // a thread's bits: rows 2q | 2q + 1 << 16
static constexpr uint32_t CT_SPARSE = 48;                           // set entries per 512-entry row
static constexpr uint32_t CT_STAGE = 2 * 512 + 4;                   // staging words per warp (two rows)
static constexpr uint32_t CT_LOCAL = 1024;                          // a sparse tile's indices fit its 4 KiB of bits
static constexpr uint32_t CT_GROUP_SHIFT = 7;                       // 128 tiles per counter group
static_assert((1u << CT_GROUP_SHIFT) <= CT_THREADS, "one tile count per thread");

/// bit k of the result = (byte k of w != 0) (k < 4).  The same predicate as the
/// single-pass kernel for small masks (any non-zero byte selects the entry), so
/// that count and indices do not depend on which path serves a call; the
/// reference defines only 0 / 1 (jit.h:2377-2379).
B200_DEVICE uint32_t pack4(uint32_t w) {
    return ((((nonzero_bytes(w) >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xfu;
}

B200_DEVICE uint32_t pack16(uint4 v) {
    return pack4(v.x) | (pack4(v.y) << 4) | (pack4(v.z) << 8) | (pack4(v.w) << 12);
}

/// Inclusive scans over the lanes of all CT_ROWS row counts at once: three 10-bit
/// fields per register (a row holds at most 512 set entries).  c[j]: set entries of
/// this lane's word of row j, P: packed inclusive scans, T: packed row totals.
struct CtScan {
    static constexpr int J = CT_ROWS, NP = (J + 2) / 3;
    uint32_t c[J], P[NP], T[NP];
    B200_DEVICE uint32_t incl(int j) const { return (P[j / 3] >> (10 * (j % 3))) & 0x3ffu; }
    B200_DEVICE uint32_t total(int j) const { return (T[j / 3] >> (10 * (j % 3))) & 0x3ffu; }
    B200_DEVICE void run(const uint32_t (&hp)[CT_ROWS / 2], uint32_t lane) {
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
        #pragma unroll
        for (int i = 0; i < NP; ++i)
            T[i] = __shfl_sync(FULL_MASK, P[i], 31);
    }
};

/// Sparse expansion of a warp's 4096 entries: every lane walks the set bits of its
/// own words and stores the entry indices (item0 = index of bit 0 of this lane's word
/// of row 0) to out[first ..] in ascending order.
B200_DEVICE void ct_walk_scanned(const uint32_t (&hp)[CT_ROWS / 2], const CtScan &sc, uint32_t first,
                                 uint32_t item0, uint32_t *__restrict__ out) {
    uint32_t row_first = first;
    #pragma unroll
    for (int j = 0; j < CT_ROWS; ++j) {
        uint32_t word = (hp[j / 2] >> (16 * (j & 1))) & 0xffffu;
        uint32_t o = row_first + sc.incl(j) - sc.c[j]; // slot of this lane's first set entry
        while (word) {
            const uint32_t b = __ffs(word) - 1;
            word &= word - 1;
            out[o++] = item0 + j * 512 + b;
        }
        row_first += sc.total(j);
    }
}

/// Virtual layout: entry i of the mask is byte (i + mis) of the 16-byte aligned
/// array (in - mis); bits / tiles are indexed in that virtual space.
__global__ void __launch_bounds__(CT_THREADS)
compress_pack_kernel(const uint8_t *__restrict__ in, uint64_t size, uint32_t mis,
                     CtBits *__restrict__ bits, uint32_t *__restrict__ counts,
                     uint32_t *__restrict__ group_counts) {
    __shared__ uint32_t s_cnt[CT_WARPS];
    __shared__ __align__(16) uint16_t s_tr[CT_WARPS][CT_ROWS * 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *vin = in - mis;
    const uint64_t vend = size + mis; // valid virtual bytes: [mis, vend)
    const uint64_t vb0 = (uint64_t) blockIdx.x * CT_TILE + warp * CT_WARP_ENTRIES + lane * 16;

    uint4 v[CT_ROWS];
    if ((uint64_t) blockIdx.x * CT_TILE >= mis && (uint64_t) (blockIdx.x + 1) * CT_TILE <= vend) {
        #pragma unroll
        for (int j = 0; j < CT_ROWS; ++j)
            v[j] = ld_stream(vin + vb0 + j * 512);
    } else {
        #pragma unroll
        for (int j = 0; j < CT_ROWS; ++j) {
            const uint64_t vb = vb0 + j * 512;
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
            v[j] = x;
        }
    }
    CtBits hp;
    uint32_t pc = 0;
    #pragma unroll
    for (int q = 0; q < CT_ROWS / 2; ++q) {
        hp.w[q] = pack16(v[2 * q]) | (pack16(v[2 * q + 1]) << 16);
        pc += __popc(hp.w[q]);
    }
    const uint32_t c = __reduce_add_sync(FULL_MASK, pc);
    if (lane == 0)
        s_cnt[warp] = c;
    __syncthreads();
    uint32_t t = 0, first = 0; // set entries of the tile / in front of this warp
    #pragma unroll
    for (int w = 0; w < CT_WARPS; ++w) {
        const uint32_t cw = s_cnt[w];
        t += cw;
        first += (uint32_t) w < warp ? cw : 0u;
    }
    if (tid == 0) {
        counts[blockIdx.x] = t;
        if (t)
            atomicAdd(group_counts + (blockIdx.x >> CT_GROUP_SHIFT), t);
    }
    if (t == 0)
        return; // (nothing will look at this tile's slot)
    if (t > CT_LOCAL) {
        bits[(size_t) blockIdx.x * CT_THREADS + tid] = hp;
        return;
    }
    // ---- sparse tile: its <= CT_LOCAL indices go, in order, where its bits would have
    // gone (the expand kernel then only moves them to their final place).  The warp's
    // 256 half words are transposed through shared memory so that lane l holds the 128
    // CONSECUTIVE entries [128 l, 128 l + 128) of the warp's 4096: one lane scan instead
    // of one per row.
    if (c == 0)
        return; // warp-uniform
    uint16_t *tr = s_tr[warp];
    #pragma unroll
    for (int j = 0; j < CT_ROWS; ++j)
        tr[j * 32 + lane] = (uint16_t) (hp.w[j / 2] >> (16 * (j & 1)));
    __syncwarp();
    const uint4 w4 = ((const uint4 *) tr)[lane];
    const uint32_t w[4] = { w4.x, w4.y, w4.z, w4.w };
    const uint32_t mine = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
    uint32_t incl = mine;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= (uint32_t) d)
            incl += up;
    }
    uint32_t *slot = (uint32_t *) (bits + (size_t) blockIdx.x * CT_THREADS) + first + incl - mine;
    const uint32_t item0 = (uint32_t) blockIdx.x * CT_TILE + warp * CT_WARP_ENTRIES + lane * 128 - mis;
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t word = w[k];
        while (word) {
            const uint32_t b = __ffs(word) - 1;
            word &= word - 1;
            *slot++ = item0 + k * 32 + b;
        }
    }
}

__global__ void __launch_bounds__(CT_THREADS)
compress_expand_kernel(const CtBits *__restrict__ bits, const uint32_t *__restrict__ counts,
                       const uint32_t *__restrict__ group_counts, uint32_t ntiles, uint32_t mis,
                       uint32_t *__restrict__ out, uint32_t *__restrict__ count_out) {
    __shared__ uint32_t s_wtot[CT_WARPS], s_before[CT_WARPS];
    extern __shared__ __align__(16) uint32_t ct_stage[]; // CT_WARPS * CT_STAGE words
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    uint32_t *stage = ct_stage + warp * CT_STAGE;

    // the tile's count, its slot (bits, or -- sparse tiles, expanded by the pack kernel
    // already -- up to CT_LOCAL indices) and the counters in front of it are requested
    // together: one memory round trip per CTA
    const uint32_t tcount = __ldg(counts + tile);
    uint32_t hp[CT_ROWS / 2];
    {
        const uint4 raw = __ldg((const uint4 *) (bits + (size_t) tile * CT_THREADS + tid));
        hp[0] = raw.x; hp[1] = raw.y; hp[2] = raw.z; hp[3] = raw.w;
    }
    static_assert(CT_ROWS == 8, "one 16-byte load of bits per thread");
    // set entries in front of this tile: whole groups of tiles, then the tiles of
    // this tile's own group
    uint32_t before = 0;
    {
        const uint32_t group = tile >> CT_GROUP_SHIFT, t0 = group << CT_GROUP_SHIFT;
        if (t0 + tid < tile) // CT_THREADS >= tiles per group
            before = __ldg(counts + t0 + tid);
        for (uint32_t g = tid; g < group; g += CT_THREADS)
            before += __ldg(group_counts + g);
        before = __reduce_add_sync(FULL_MASK, before);
    }

    // tiles without set entries have nothing to do (the last tile leaves the count)
    if (tcount == 0 && tile != ntiles - 1)
        return;
    const bool local = tcount <= CT_LOCAL;
    constexpr int J = CT_ROWS;
    CtScan sc;
    uint32_t wtotal = 0;
    if (!local) {
        sc.run(hp, lane);
        #pragma unroll
        for (int j = 0; j < J; ++j)
            wtotal += sc.total(j);
    }
    if (lane == 0) {
        s_wtot[warp] = wtotal;
        s_before[warp] = before;
    }
    __syncthreads();
    uint32_t first = 0, tile_total = 0;
    #pragma unroll
    for (int w = 0; w < CT_WARPS; ++w) {
        first += s_before[w] + ((uint32_t) w < warp ? s_wtot[w] : 0u);
        tile_total += s_wtot[w];
    }
    if (local) { // (block-uniform) first = everything in front of the tile
        if (tile == ntiles - 1 && tid == 0)
            *count_out = first + tcount;
        #pragma unroll
        for (uint32_t k = 0; k < 4; ++k)
            if (4 * tid + k < tcount)
                out[first + 4 * tid + k] = hp[k];
        return;
    }
    if (tile == ntiles - 1 && tid == 0)
        *count_out = first + tile_total; // warp 0: first = everything in front of the tile

    // entry index of bit 0 of this lane's word of row 0 (may wrap below zero for
    // the masked-out head bytes, which are never set)
    const uint32_t item0 = tile * CT_TILE + warp * CT_WARP_ENTRIES + lane * 16 - mis;
    if (wtotal == 0)
        return; // warp-uniform
    if (wtotal <= CT_SPARSE * J) {
        // sparse: every lane walks its own set bits and stores straight to global
        ct_walk_scanned(hp, sc, first, item0, out);
        return;
    }

    // dense: the indices of two rows at a time (up to 1024) are staged contiguously;
    // stage[a + i] <-> out[pair_first + i], where 'a' makes 16-byte aligned staging
    // words coincide with 16-byte aligned global addresses
    uint32_t pair_first = first;
    #pragma unroll
    for (int q = 0; q < J / 2; ++q) {
        const uint32_t t0 = sc.total(2 * q);
        const uint32_t t1 = sc.total(2 * q + 1);
        const uint32_t pair_total = t0 + t1;
        if (pair_total == 0)
            continue; // warp-uniform
        uint32_t *dst = out + pair_first;
        const uint32_t a = (uint32_t) (((uintptr_t) dst >> 2) & 3u);
        if (q > 0) { // the previous pair's bulk copy may still be reading the staging area
            if (lane == 0)
                bulk_wait_read<0>();
            __syncwarp();
        }
        // nearly full pairs take the rotated order: with every lane starting at bit 0
        // the lanes of a full row hit the staging area 16 words apart, a 16-way bank
        // conflict (ncu at density 0.99: 52 M of 67 M shared wavefronts were
        // conflicts); starting lane l at bit (l / 2) % 16 spreads them over all banks
        // at the price of three more instructions per bit
        const bool rotated = pair_total > 2 * 400;
        #pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = 2 * q + r;
            const uint32_t incl = sc.incl(j);
            const uint32_t word = (hp[j / 2] >> (16 * (j & 1))) & 0xffffu;
            const uint32_t lane_base =
                (uint32_t) __cvta_generic_to_shared(stage + a + (r ? t0 : 0u) + (incl - sc.c[j]));
            if (!rotated) {
                uint32_t sp = lane_base, v = item0 + j * 512;
                #pragma unroll
                for (int b = 0; b < 16; ++b) {
                    // predicated store + predicated pointer bump, unconditional value bump
                    asm volatile("{\n\t.reg .pred p;\n\t"
                                 "setp.ne.u32 p, %2, 0;\n\t"
                                 "@p st.shared.u32 [%0], %1;\n\t"
                                 "@p add.u32 %0, %0, 4;\n\t}"
                                 : "+r"(sp) : "r"(v), "r"(word & (1u << b)) : "memory");
                    v++;
                }
            } else {
                const uint32_t rot = (lane >> 1) & 15u, wrap = 16u - rot;
                const uint32_t wr = ((word >> rot) | (word << wrap)) & 0xffffu;
                uint32_t sp = lane_base + 4u * __popc(word & ((1u << rot) - 1u));
                uint32_t v = item0 + j * 512 + rot;
                #pragma unroll
                for (uint32_t b = 0; b < 16; ++b) {
                    if (b == wrap) { // bits [0, rot) follow
                        sp = lane_base;
                        v -= 16;
                    }
                    asm volatile("{\n\t.reg .pred p;\n\t"
                                 "setp.ne.u32 p, %2, 0;\n\t"
                                 "@p st.shared.u32 [%0], %1;\n\t"
                                 "@p add.u32 %0, %0, 4;\n\t}"
                                 : "+r"(sp) : "r"(v), "r"(wr & (1u << b)) : "memory");
                    v++;
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        const uint32_t end = a + pair_total;          // staging words [a, end) are valid
        const uint32_t lo = (a + 3u) & ~3u, hi = end & ~3u;
        if (hi > lo) {
            if (lane == 0) {
                bulk_s2g(dst - a + lo, stage + lo, (hi - lo) * 4);
                bulk_commit();
            }
            // ragged ends: [a, lo) and [hi, end)
            if (lane < lo - a)
                dst[lane] = stage[a + lane];
            else if (lane >= 4 && lane - 4 < end - hi)
                dst[hi - a + lane - 4] = stage[hi + lane - 4];
        } else {
            for (uint32_t i = lane; i < pair_total; i += 32)
                dst[i] = stage[a + i];
        }
        pair_first += pair_total;
    }
    // the staging area must outlive the bulk copy that reads it
    if (lane == 0)
        bulk_wait_read<0>();
}

static int compress_tiles(cudaStream_t stream, const uint8_t *in, uint64_t size, uint32_t *out,
                          uint32_t *count_dev) {
    const uint32_t mis = (uint32_t) ((uintptr_t) in & 15);
    const uint32_t ntiles = (uint32_t) ceil_div(size + mis, CT_TILE);
    const size_t bits_bytes = (size_t) ntiles * CT_THREADS * sizeof(CtBits);
    const uint32_t ngroups = (ntiles >> CT_GROUP_SHIFT) + 1;
    uint8_t *scratch = (uint8_t *) temp_alloc(bits_bytes + ((size_t) ntiles + ngroups) * 4, stream);
    if (!scratch)
        return fail(B200_ERR_CUDA, "jit_compress(): out of memory (%zu bytes)", bits_bytes);
    CtBits *bits = (CtBits *) scratch;
    uint32_t *counts = (uint32_t *) (scratch + bits_bytes);
    uint32_t *group_counts = counts + ntiles;
    cudaError_t err = cudaMemsetAsync(group_counts, 0, (size_t) ngroups * 4, stream);
    if (err != cudaSuccess) {
        temp_free(scratch, stream);
        return cuda_fail(err, "cudaMemsetAsync");
    }
    compress_pack_kernel<<<ntiles, CT_THREADS, 0, stream>>>(in, size, mis, bits, counts, group_counts);
    constexpr size_t stage_bytes = (size_t) CT_WARPS * CT_STAGE * 4;
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !configured[dev].load(std::memory_order_relaxed)) {
        err = cudaFuncSetAttribute(compress_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int) stage_bytes);
        if (err != cudaSuccess) {
            temp_free(scratch, stream);
            return cuda_fail(err, "cudaFuncSetAttribute");
        }
        if (dev < 64)
            configured[dev].store(true, std::memory_order_relaxed);
    }
    compress_expand_kernel<<<ntiles, CT_THREADS, stage_bytes, stream>>>(bits, counts, group_counts, ntiles,
                                                                        mis, out, count_dev);
    count_launch(1);
    temp_free(scratch, stream);
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
    HistoryScope hs(stream, B200_KERNEL_COMPRESS, size);
    // small masks: one launch of the single-pass kernel; large ones: bit-packed tiles
    if (size > 32768)
        return compress_tiles(stream, in, size, out, count_dev);

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
    if ((rc = sync_forbidden())) // the reference calls jitc_sync_thread (src/cuda_ts.cpp:759)
        return rc;
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
