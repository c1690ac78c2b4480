// mkperm.cu -- bucketing permutation behind vectorised method dispatch.
//
// Replaces CUDAThreadState::block_mkperm (src/cuda_ts.cpp:788-975), the batched
// transpose (:765-786) and resources/mkperm.cuh (phase_1 / phase_3 / phase_4 in
// their tiny / small / large variants).
//
// Structure (counting sort, one or more stable digit passes):
//   rows     every sorting group is cut into S contiguous rows; ONE WARP owns a
//            row and walks it in index order, 32 keys per step.  A warp keeps
//            a private shared-memory table of B counters (histogram pass) or
//            running output offsets (placement pass), so no atomics are needed
//            and the permutation is STABLE for every bucket count (the
//            reference is stable only in its "tiny" variant).
//   counts   the histogram pass writes counts bucket-major ([group][bucket][row])
//            so that ONE segmented exclusive scan (scan.cu, block size
//            B * S) yields every row's first output slot -- the reference needs
//            two batched transposes around its scan.
//   place    per step: match.any finds the lanes with the same key, the lowest
//            such lane advances the row's counter by the group size, everyone
//            stores its index at counter + rank.
//   digits   more than 2048 buckets do not fit a per-warp table; the key is
//            then split into <= 11-bit digits sorted least-significant first
//            (each pass stable => the result is the same stable permutation).
//            The reference instead falls back to global atomics into 148
//            private global histograms.
//   offsets  non-empty buckets are emitted as (id, start, size, 0) records in
//            ascending id order by a single-CTA compaction.
#include "common.cuh"
#include "pipeline.cuh"

#include <algorithm>
#include <cstdlib>

#include <nvtx3/nvToolsExt.h>

namespace b200 {

/// NVTX range (header-only NVTX v3: a no-op unless a profiler is attached) -- the
/// reference brackets jit_block_mkperm with a ProfilerPhase (src/util.cpp:218-229)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static constexpr int MKPERM_WARPS = 16;
static constexpr int MKPERM_THREADS = MKPERM_WARPS * 32;
static constexpr uint32_t MKPERM_MAX_BINS = 2048; // 16 warps * 2048 * 4 B = 128 KiB

struct MkpermGeom {
    uint64_t size;        // total number of keys
    uint64_t group_size;  // block_size of the API
    uint64_t group0;      // first group handled by this launch
    uint32_t ngroups;     // groups handled by this launch
    uint32_t rows_per_group;
    uint64_t row_len;
    uint32_t bins;        // counters per row in this pass
    uint32_t shift;       // digit = (key >> shift) & mask
    uint32_t mask;
};

B200_DEVICE uint32_t digit_of(uint32_t key, const MkpermGeom &g) {
    uint32_t d = (key >> g.shift) & g.mask;
    return min(d, g.bins - 1); // out-of-range keys must not corrupt shared memory
}

B200_DEVICE bool row_range(const MkpermGeom &g, uint64_t row, uint64_t &start, uint64_t &end,
                           uint32_t &grp, uint32_t &slice) {
    uint64_t nrows = (uint64_t) g.ngroups * g.rows_per_group;
    if (row >= nrows)
        return false;
    grp = (uint32_t) (row / g.rows_per_group);
    slice = (uint32_t) (row - (uint64_t) grp * g.rows_per_group);
    uint64_t gstart = (g.group0 + grp) * g.group_size;
    start = gstart + (uint64_t) slice * g.row_len;
    end = min(min(start + g.row_len, gstart + g.group_size), g.size);
    if (start > end)
        start = end;
    return true;
}

/// Histogram pass: counts[(grp * bins + b) * rows_per_group + slice]
__global__ void __launch_bounds__(MKPERM_THREADS)
mkperm_hist_kernel(const uint32_t *__restrict__ keys, uint32_t *__restrict__ counts,
                   const MkpermGeom g) {
    extern __shared__ uint32_t mk_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *h = mk_smem + (size_t) warp * g.bins;

    uint64_t start, end;
    uint32_t grp, slice;
    if (!row_range(g, (uint64_t) blockIdx.x * MKPERM_WARPS + warp, start, end, grp, slice))
        return;

    for (uint32_t b = lane; b < g.bins; b += 32)
        h[b] = 0;
    __syncwarp();

    constexpr int U = 8;
    uint64_t k = start + lane;
    for (; k + (U - 1) * 32 < end; k += U * 32) {
        uint32_t key[U];
        #pragma unroll
        for (int u = 0; u < U; ++u)
            key[u] = __ldg(keys + k + u * 32);
        #pragma unroll
        for (int u = 0; u < U; ++u)
            atomicAdd(&h[digit_of(key[u], g)], 1u);
    }
    for (; k < end; k += 32)
        atomicAdd(&h[digit_of(__ldg(keys + k), g)], 1u);
    __syncwarp();

    uint32_t *dst = counts + ((uint64_t) grp * g.bins) * g.rows_per_group + slice;
    for (uint32_t b = lane; b < g.bins; b += 32)
        dst[(uint64_t) b * g.rows_per_group] = h[b];
}

/// Placement pass.  'offsets' are the exclusively scanned counts.  The element
/// written is idx_in[k] (or k itself when idx_in == NULL); keys_out (optional)
/// receives the key at the same slot for the next digit pass.
__global__ void __launch_bounds__(MKPERM_THREADS)
mkperm_place_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ idx_in,
                    const uint32_t *__restrict__ offsets, uint32_t *__restrict__ perm_out,
                    uint32_t *__restrict__ keys_out, const MkpermGeom g) {
    extern __shared__ uint32_t mk_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    uint32_t *h = mk_smem + (size_t) warp * g.bins;

    uint64_t start, end;
    uint32_t grp, slice;
    if (!row_range(g, (uint64_t) blockIdx.x * MKPERM_WARPS + warp, start, end, grp, slice))
        return;

    const uint32_t gbase = (uint32_t) ((g.group0 + grp) * g.group_size);
    const uint32_t *src = offsets + ((uint64_t) grp * g.bins) * g.rows_per_group + slice;
    for (uint32_t b = lane; b < g.bins; b += 32)
        h[b] = src[(uint64_t) b * g.rows_per_group] + gbase;
    __syncwarp();

    constexpr int U = 8;
    for (uint64_t k0 = start; k0 < end; k0 += U * 32) {
        uint32_t key[U], idx[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t k = k0 + u * 32 + lane;
            bool valid = k < end;
            key[u] = valid ? __ldg(keys + k) : 0;
            idx[u] = valid ? (idx_in ? __ldg(idx_in + k) : (uint32_t) k) : 0;
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t k = k0 + u * 32 + lane;
            if (k0 + u * 32 >= end)
                break; // warp-uniform
            bool valid = k < end;
            uint32_t active = __ballot_sync(FULL_MASK, valid);
            if (valid) {
                uint32_t d = digit_of(key[u], g);
                uint32_t peers = __match_any_sync(active, d);
                uint32_t leader = __ffs(peers) - 1;
                uint32_t rank = __popc(peers & lt_mask);
                uint32_t base = 0;
                if (lane == leader) {
                    base = h[d];
                    h[d] = base + __popc(peers);
                }
                base = __shfl_sync(peers, base, leader);
                uint32_t pos = base + rank;
                perm_out[pos] = idx[u];
                if (keys_out)
                    keys_out[pos] = key[u];
            }
            __syncwarp();
        }
    }
}

/// Whole-array histogram of the keys in [lo, lo + bins) with a per-CTA
/// shared-memory table that is flushed with global atomics (bins * 4 bytes must
/// fit the dynamic shared memory).  Keys >= limit count as limit - 1 (out-of-range
/// keys must not corrupt memory); keys outside the sub-range are skipped.
__global__ void __launch_bounds__(1024)
histogram_smem_kernel(const uint32_t *__restrict__ keys, uint64_t size, uint32_t lo, uint32_t bins,
                      uint32_t limit, uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t mk_smem[];
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
        mk_smem[b] = 0;
    __syncthreads();
    uint64_t chunk = (size + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 3) & ~3ull;
    uint64_t start = min((uint64_t) blockIdx.x * chunk, size), end = min(start + chunk, size);
    auto add = [&](uint32_t k) {
        const uint32_t b = min(k, limit - 1) - lo; // wraps for keys below lo
        if (b < bins)
            atomicAdd(&mk_smem[b], 1u);
    };
    bool aligned = ((uintptr_t) keys & 15) == 0;
    if (aligned) {
        uint64_t nvec = (end - start) / 4;
        const uint4 *v = (const uint4 *) (keys + start);
        uint64_t q = threadIdx.x;
        for (; q + blockDim.x < nvec; q += 2 * blockDim.x) {
            uint4 k4 = ld_stream(v + q), k5 = ld_stream(v + q + blockDim.x);
            add(k4.x); add(k4.y); add(k4.z); add(k4.w);
            add(k5.x); add(k5.y); add(k5.z); add(k5.w);
        }
        for (; q < nvec; q += blockDim.x) {
            uint4 k4 = ld_stream(v + q);
            add(k4.x); add(k4.y); add(k4.z); add(k4.w);
        }
        start += nvec * 4;
    }
    for (uint64_t k = start + threadIdx.x; k < end; k += blockDim.x)
        add(keys[k]);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) {
        uint32_t c = mk_smem[b];
        if (c)
            atomicAdd(&hist[lo + b], c);
    }
}

/// Whole-array histogram straight into global memory (very large bucket counts)
__global__ void __launch_bounds__(256)
histogram_global_kernel(const uint32_t *__restrict__ keys, uint64_t size, uint32_t bins,
                        uint32_t *__restrict__ hist) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < size; k += stride)
        atomicAdd(&hist[min(__ldg(keys + k), bins - 1)], 1u);
}

/// Emit (id, start, size, 0) for every non-empty bucket in ascending id order.
/// starts[b * stride] is the first output slot of bucket b; 'records' has room
/// for 4 * bins + 1 words, the last one receiving the number of records.  One
/// CTA: a warp owns ceil(bins / 1024) rows of 32 consecutive buckets (coalesced
/// loads), counts its non-empty buckets, one block scan places the warps, a
/// second walk writes the records.
__global__ void __launch_bounds__(1024)
mkperm_offsets_kernel(const uint32_t *__restrict__ starts, uint64_t stride, uint32_t bins,
                      uint32_t size, uint32_t *__restrict__ records) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rows = (bins + 1023u) / 1024u;
    const uint32_t warp_base = warp * rows * 32u;
    const uint32_t lt = (1u << lane) - 1u;

    auto row = [&](uint32_t j, uint32_t &b, uint32_t &st, uint32_t &nx) -> uint32_t { // ballot of non-empty
        b = warp_base + j * 32u + lane;
        st = b < bins ? __ldg(starts + (uint64_t) b * stride) : size;
        nx = __shfl_down_sync(FULL_MASK, st, 1);
        if (lane == 31)
            nx = b + 1 < bins ? __ldg(starts + (uint64_t) (b + 1) * stride) : size;
        return __ballot_sync(FULL_MASK, b < bins && nx != st);
    };

    uint32_t mine = 0; // warp-uniform
    #pragma unroll 4
    for (uint32_t j = 0; j < rows; ++j) {
        uint32_t b, st, nx;
        mine += __popc(row(j, b, st, nx));
    }
    if (lane == 0)
        warp_sums[warp] = mine;
    __syncthreads();
    uint32_t wsum = warp_sums[lane], wincl = wsum;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t up = __shfl_up_sync(FULL_MASK, wincl, d);
        if (lane >= (uint32_t) d)
            wincl += up;
    }
    uint32_t rec = __shfl_sync(FULL_MASK, wincl - wsum, warp);
    const uint32_t total = __shfl_sync(FULL_MASK, wincl, 31);
    #pragma unroll 4
    for (uint32_t j = 0; j < rows; ++j) {
        uint32_t b, st, nx;
        const uint32_t flags = row(j, b, st, nx);
        if ((flags >> lane) & 1u)
            ((uint4 *) records)[rec + __popc(flags & lt)] = make_uint4(b, st, nx - st, 0);
        rec += __popc(flags);
    }
    if (tid == 0)
        records[4 * (size_t) bins] = total;
}

/// The records of mkperm_offsets_kernel re-ordered by bucket size, largest first (ties:
/// ascending bucket id) -- the order jitc_var_call_reduce launches the callees in
/// (src/call.cpp:1346-1352 sorts on the host after the event wait).  One CTA: bitonic
/// sort of 64-bit keys {size, ~record index} in shared memory (up to RS_SMEM records) or
/// in `scratch` (next_pow2(bins) keys) beyond that.
static constexpr uint32_t RS_SMEM = 4096;
__global__ void __launch_bounds__(1024)
mkperm_sort_records_kernel(const uint32_t *__restrict__ records, uint32_t bins,
                           uint32_t *__restrict__ sorted, unsigned long long *__restrict__ scratch) {
    __shared__ unsigned long long s_keys[RS_SMEM];
    const uint32_t tid = threadIdx.x;
    const uint32_t count = records[4 * (size_t) bins];
    uint32_t padded = 1;
    while (padded < count)
        padded <<= 1;
    unsigned long long *keys = padded <= RS_SMEM ? s_keys : scratch;
    for (uint32_t i = tid; i < padded; i += 1024)
        keys[i] = i < count ? ((unsigned long long) records[4 * (size_t) i + 2] << 32) | (0xffffffffu - i) : 0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= padded; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < padded / 2; t += 1024) {
                // the t-th pair of this step: i has bit j clear
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                const unsigned long long a = keys[i], b = keys[p];
                const bool descending = (i & k) == 0;
                if (descending ? a < b : a > b) {
                    keys[i] = b;
                    keys[p] = a;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < count; i += 1024) {
        const uint32_t src = 0xffffffffu - (uint32_t) keys[i];
        ((uint4 *) sorted)[i] = ((const uint4 *) records)[src];
    }
    if (tid == 0)
        sorted[4 * (size_t) bins] = count;
}

/// Copy the records (4 * bins + 1 words, device) into the caller's host-accessible
/// `offsets`; by_size: ordered by bucket size first.
static int deliver_records(cudaStream_t stream, const uint32_t *records, uint32_t bins, uint32_t *offsets,
                           bool by_size) {
    const size_t words = (size_t) bins * 4 + 1;
    if (!by_size)
        return cuda_fail(cudaMemcpyAsync(offsets, records, words * 4, cudaMemcpyDeviceToHost, stream),
                         "cudaMemcpyAsync(offsets)");
    uint32_t padded = 1;
    while (padded < bins)
        padded <<= 1;
    const size_t sorted_words = (words + 3) & ~(size_t) 3;
    uint32_t *sorted = (uint32_t *) temp_alloc(sorted_words * 4 + (padded > RS_SMEM ? (size_t) padded * 8 : 0), stream);
    if (!sorted)
        return fail(B200_ERR_CUDA, "jit_var_call_reduce(): out of memory");
    mkperm_sort_records_kernel<<<1, 1024, 0, stream>>>(records, bins, sorted,
                                                       (unsigned long long *) (sorted + sorted_words));
    count_launch();
    int rc = cuda_fail(cudaMemcpyAsync(offsets, sorted, words * 4, cudaMemcpyDeviceToHost, stream),
                       "cudaMemcpyAsync(offsets)");
    temp_free(sorted, stream);
    return rc;
}


/// starts[b] = number of keys below b = first slot of bucket b in the finished
/// permutation, found by binary search over the permutation itself (keys are
/// sorted along it).  For many buckets this beats a histogram of the keys, which
/// no longer fits one shared-memory table: 26 dependent steps of two gathers per
/// bucket against another one or two sweeps over all keys.
/// Sixteen lanes per bucket probe sixteen pivots of the remaining range at once
/// (17-ary search: 7 dependent steps for 2^26 keys instead of 26).
__global__ void __launch_bounds__(256)
mkperm_bounds_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ perm,
                     uint32_t size, uint32_t index_base, uint32_t bins, uint32_t *__restrict__ starts) {
    const uint32_t sub = threadIdx.x & 15, lane = threadIdx.x & 31;
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const uint32_t group = 0xffffu << (lane & 16); // the lanes of this bucket
    const bool valid = b < bins; // (uniform per group of 16 lanes; all lanes take part in the shuffles)
    uint32_t lo = 0, hi = valid ? size : 0; // first position whose key is >= b lies in [lo, hi]
    auto key_at = [&](uint32_t pos) {
        return min(__ldg(keys + (__ldg(perm + pos) - index_base)), bins - 1);
    };
    while (__any_sync(FULL_MASK, hi - lo > 16)) {
        const uint32_t len = hi - lo;
        if (len > 16) {
            // pivots lo + ceil((j + 1) * len / 17) - 1 < hi, ascending in j = 0 .. 15
            const uint32_t piv = lo + (uint32_t) (((uint64_t) (sub + 1) * len + 16) / 17) - 1;
            const uint32_t ge = __ballot_sync(group, key_at(piv) >= b) & group;
            // first pivot whose key is >= b bounds the answer from above, the one before from below
            const uint32_t j = ge ? (uint32_t) __ffs(ge) - 1 - (lane & 16) : 16;
            const uint32_t piv_hi = __shfl_sync(group, piv, (lane & 16) + min(j, 15u));
            const uint32_t piv_lo = __shfl_sync(group, piv, (lane & 16) + (j ? j - 1 : 0));
            if (j < 16)
                hi = piv_hi;
            if (j > 0)
                lo = piv_lo + 1;
        }
    }
    // at most 16 candidates left: [lo, hi)
    const uint32_t pos = lo + sub;
    const bool ge = pos >= hi || key_at(pos) >= b;
    const uint32_t m = __ballot_sync(group, ge) & group;
    if (valid && sub == 0)
        starts[b] = m ? lo + (uint32_t) __ffs(m) - 1 - (lane & 16) : hi;
}

/// mkperm_offsets_kernel for many buckets, in two launches of one CTA per 1024 buckets:
/// `count` = non-empty buckets per CTA, `emit` sums the counts in front of it and writes
/// the records (ascending id).  starts[b] = first output slot of bucket b.
__global__ void __launch_bounds__(1024)
mkperm_offsets_count_kernel(const uint32_t *__restrict__ starts, uint32_t bins, uint32_t size,
                            uint32_t *__restrict__ block_counts) {
    const uint32_t b = blockIdx.x * 1024 + threadIdx.x;
    bool nonempty = false;
    if (b < bins)
        nonempty = (b + 1 < bins ? __ldg(starts + b + 1) : size) != __ldg(starts + b);
    const uint32_t c = (uint32_t) __syncthreads_count(nonempty);
    if (threadIdx.x == 0)
        block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
mkperm_offsets_emit_kernel(const uint32_t *__restrict__ starts, uint32_t bins, uint32_t size,
                           const uint32_t *__restrict__ block_counts, uint32_t *__restrict__ records) {
    __shared__ uint32_t s_part[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // records in front of this CTA (and, for the last CTA, in total)
    uint32_t before = 0;
    for (uint32_t i = tid; i < blockIdx.x; i += 1024)
        before += __ldg(block_counts + i);
    before = __reduce_add_sync(FULL_MASK, before);
    if (lane == 0)
        s_part[warp] = before;
    __syncthreads();
    before = 0;
    #pragma unroll
    for (int w = 0; w < 32; ++w)
        before += s_part[w];
    __syncthreads();
    const uint32_t b = blockIdx.x * 1024 + tid;
    uint32_t st = 0, nx = 0;
    if (b < bins) {
        st = __ldg(starts + b);
        nx = b + 1 < bins ? __ldg(starts + b + 1) : size;
    }
    const bool nonempty = b < bins && nx != st;
    const uint32_t flags = __ballot_sync(FULL_MASK, nonempty);
    if (lane == 0)
        s_part[warp] = __popc(flags);
    __syncthreads();
    uint32_t rec = before, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < 32; ++w) {
        const uint32_t c = s_part[w];
        rec += w < warp ? c : 0u;
        total += c;
    }
    if (nonempty)
        ((uint4 *) records)[rec + __popc(flags & ((1u << lane) - 1u))] = make_uint4(b, st, nx - st, 0);
    if (blockIdx.x == gridDim.x - 1 && tid == 0)
        records[4 * (size_t) bins] = before + total;
}

/// Records from contiguous bucket starts (starts[b], b < bins): one CTA up to 8192
/// buckets, the two-launch form beyond.  `scratch`: ceil(bins / 1024) words.
static void offsets_from_starts(cudaStream_t stream, const uint32_t *starts, uint32_t bins, uint32_t size,
                                uint32_t *records, uint32_t *scratch) {
    if (bins <= 8192) {
        mkperm_offsets_kernel<<<1, 1024, 0, stream>>>(starts, 1, bins, size, records);
        count_launch();
        return;
    }
    const uint32_t grid = (uint32_t) ceil_div(bins, 1024);
    mkperm_offsets_count_kernel<<<grid, 1024, 0, stream>>>(starts, bins, size, scratch);
    mkperm_offsets_emit_kernel<<<grid, 1024, 0, stream>>>(starts, bins, size, scratch, records);
    count_launch(2);
}

// ------------------------------------------------ single sorting group: tiles
//
// block_size == size (what vectorised method dispatch passes) and every block_size
// of at least half a tile: the array is cut into tiles of 8192 keys that never cross
// a sorting group, a tile into two HALVES of 4096 keys (the keys of the lower / upper
// half of the ranking kernel's threads).
//   tile_hist   per-half digit counts, written digit-major (table[d][tile][half]) so
//               that ONE exclusive scan of the flattened table (scan_fast.cu)
//               yields the first output slot of every (digit, tile, half) triple;
//   rank_place  re-reads the tile, ranks its keys stably, sorts the tile by digit
//               in shared memory and writes every digit's two runs with bulk copies
//               (further down).  The per-row variant above has every warp
//               trickle single elements into `bins` output streams at once,
//               which thrashes L2 (measured on B200: 6.5x DRAM write
//               amplification, 2-4 ms for 2^26 keys).
// Any radix pass has to see ALL keys before it can place the first one (the first
// slot of bucket 1 is the number of keys in bucket 0), so with keys far beyond the
// on-chip memories a pass moves at least 4 (count) + 4 (rank) + 4 (write) bytes per
// key: 12 B/key, not the 8 B/key of the permutation's inputs and outputs.
static constexpr int MT_THREADS = 512;
static constexpr uint32_t MT_MAX_BINS = 64;                // 6-bit digits
static constexpr uint32_t MT_GROUP = 16;                   // half tiles counted by one CTA

/// Tiles of the ranked path never cross a sorting group: group g owns the tiles
/// [g * tpg, (g + 1) * tpg), the last tile of a group may be short.  One group:
/// group_size = size, tpg = number of tiles.
struct TileGeom {
    uint32_t size;        // keys in total
    uint32_t group_size;  // keys per sorting group
    uint32_t tpg;         // tiles per group
    uint32_t ntiles;      // tiles in total
};

/// First key and number of keys of a tile of TILE keys
template <uint32_t TILE>
B200_DEVICE void tile_range(const TileGeom &g, uint32_t tile, uint32_t &base, uint32_t &count) {
    if (g.tpg == g.ntiles) { // one group: no division
        base = tile * TILE;
        count = min(TILE, g.size - base);
        return;
    }
    const uint32_t grp = tile / g.tpg, tg = tile - grp * g.tpg;
    const uint32_t gstart = grp * g.group_size; // (the last group starts below 2^32)
    const uint32_t glen = min(g.group_size, g.size - gstart);
    const uint32_t off = tg * TILE;
    base = gstart + off;
    count = off < glen ? min(TILE, glen - off) : 0u;
}

/// Position of (tile, half 0, digit) in the count table: [group][digit][tile of the
/// group][half], so that an exclusive scan in blocks of bins * tpg * 2 entries yields
/// the first output slot of every (digit, tile, half) triple relative to its group.
B200_DEVICE uint64_t table_index(const TileGeom &g, uint32_t tile, uint32_t digit, uint32_t bins) {
    if (g.tpg == g.ntiles)
        return ((uint64_t) digit * g.ntiles + tile) * 2;
    const uint32_t grp = tile / g.tpg, tg = tile - grp * g.tpg;
    return (((uint64_t) grp * bins + digit) * g.tpg + tg) * 2;
}

/// How the elements reach a pass / leave it
enum { RK_RAW1 = 0,   // keys = the caller's values, the only pass (-> RK_FINAL)
       RK_RAW = 1,    // keys = the caller's values, first of several passes
       RK_PAIRS = 2,  // (key, index) pairs
       RK_PACKED = 3  // one word: remaining key bits << ib | index
};
enum { RK_FINAL = 0, RK_OUT_PAIRS = 1, RK_OUT_PACKED = 2 };

/// Digit of an element: keys are clamped to kmax = bucket_count - 1 where they enter
/// (out-of-range keys must not corrupt memory), so every digit is below its `bins`.
template <int IN> B200_DEVICE uint32_t rk_digit(uint32_t x, uint32_t shift, uint32_t mask, uint32_t kmax) {
    if constexpr (IN == RK_RAW1)
        return min(x, kmax);
    else if constexpr (IN == RK_RAW)
        return (min(x, kmax) >> shift) & mask;
    else
        return (x >> shift) & mask;
}

/// table[table_index(tile, d) + half] = number of keys of that half tile whose digit
/// is d.  A CTA counts MT_GROUP consecutive half tiles into one shared-memory histogram
/// each, so that the MT_GROUP entries of a digit are contiguous in the digit-major table
/// (one 64-byte run instead of MT_GROUP scattered words).
/// Dynamic shared memory: MT_GROUP * bins counters.
template <int IN, uint32_t TILE>
__global__ void __launch_bounds__(MT_THREADS)
mkperm_tile_hist_kernel(const uint32_t *__restrict__ keys, const TileGeom g,
                        uint32_t shift, uint32_t mask, uint32_t kmax, uint32_t bins,
                        uint32_t halves, uint32_t *__restrict__ table) {
    constexpr uint32_t HALFK = TILE / 2;
    // halves == 1 (wide-digit ranking): the two halves of a tile share a histogram and
    // the table is [group][digit][tile]
    const uint32_t hshift = halves == 2 ? 0u : 1u;
    constexpr int V = HALFK / 4 / MT_THREADS, U = 8 / V; // 16-byte loads per half, halves per step
    static_assert(V >= 1 && MT_GROUP % U == 0, "eight 16-byte loads in flight per thread");
    extern __shared__ uint32_t mk_smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t tile0 = blockIdx.x * (MT_GROUP / 2);
    for (uint32_t i = tid; i < MT_GROUP * bins; i += MT_THREADS)
        mk_smem[i] = 0;
    __syncthreads();
    #pragma unroll 1
    for (uint32_t t = 0; t < MT_GROUP; t += U) {
        uint4 k4[U][V];
        bool full[U];
        uint32_t base[U], count[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            base[u] = 0;
            count[u] = 0;
            const uint32_t tile = tile0 + (t + u) / 2, half = (t + u) & 1;
            if (tile < g.ntiles) {
                uint32_t tc;
                tile_range<TILE>(g, tile, base[u], tc);
                base[u] += half * HALFK;
                count[u] = tc > half * HALFK ? min(tc - half * HALFK, HALFK) : 0u;
            }
            full[u] = count[u] == HALFK && ((uintptr_t) (keys + base[u]) & 15) == 0;
            if (full[u]) {
                const uint4 *v = (const uint4 *) (keys + base[u]);
                #pragma unroll
                for (int j = 0; j < V; ++j)
                    k4[u][j] = ld_stream(v + j * MT_THREADS + tid);
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t *h = mk_smem + ((t + u) >> hshift) * bins;
            if (full[u]) {
                #pragma unroll
                for (int j = 0; j < V; ++j) {
                    atomicAdd(&h[rk_digit<IN>(k4[u][j].x, shift, mask, kmax)], 1u);
                    atomicAdd(&h[rk_digit<IN>(k4[u][j].y, shift, mask, kmax)], 1u);
                    atomicAdd(&h[rk_digit<IN>(k4[u][j].z, shift, mask, kmax)], 1u);
                    atomicAdd(&h[rk_digit<IN>(k4[u][j].w, shift, mask, kmax)], 1u);
                }
            } else if (count[u]) {
                for (uint32_t i = tid; i < count[u]; i += MT_THREADS)
                    atomicAdd(&h[rk_digit<IN>(__ldg(keys + base[u] + i), shift, mask, kmax)], 1u);
            }
        }
    }
    __syncthreads();
    const uint32_t nslots = MT_GROUP >> hshift;
    for (uint32_t i = tid; i < nslots * bins; i += MT_THREADS) {
        const uint32_t d = i / nslots, t = i % nslots;
        const uint32_t tile = tile0 + ((t << hshift) >> 1);
        if (tile < g.ntiles) {
            const uint64_t at = table_index(g, tile, d, bins);
            table[hshift ? at >> 1 : at + (t & 1)] = mk_smem[t * bins + d];
        }
    }
}

// ------------------------------------------- single sorting group: ranked tiles
//
// match.any costs 256 issue cycles per warp instruction on B200 whenever the lanes
// disagree (tools/ubench_warp_ops.cu; a shuffle: 5, a conflict-free atoms.add: 4),
// so the kernels below rank WITHOUT any warp-wide matching, the way block radix
// ranking is classically done -- with the counters driven by shared-memory ATOMICS:
//   - a thread owns 16 CONSECUTIVE keys of the tile (blocked arrangement) and a
//     private column of 16-bit counters, one per digit.  Counting is one
//     fire-and-forget atoms.add per key on the thread's own counter (no conflicts:
//     bank = lane; no load -> add -> store chain);
//   - two counters (the same digit of thread t and of thread t + THREADS / 2) share a
//     32-bit word, the words are laid out digit-major / thread-minor, so ONE packed
//     block-wide exclusive scan of NB * THREADS / 2 words turns every counter into the
//     first STAGING slot of its (digit, thread) pair (sorted-tile slot + the run's
//     alignment shift, see below);
//   - placement is one atoms.add WITH return per key: it hands out the key's slot and
//     advances the thread's cursor in one instruction;
//   - counters cost shared memory per digit AND thread, which limits a pass to
//     6-bit digits (64 bins): more buckets are sorted least-significant digit
//     first in ceil(bits / 6) stable passes.  Between passes an element travels as
//     ONE 32-bit word (remaining key bits << index bits | index) as soon as that
//     fits, else as a (key, index) pair.
// The permutation is stable for every bucket count.
static constexpr int RK_ITEMS = 16;
static constexpr uint32_t RK_PAD = 8;  // staging slack per run (alignment)

template <int BITS, int THREADS> constexpr uint32_t rk_stage_words() {
    return THREADS * RK_ITEMS + RK_PAD * 2 * (1u << BITS);
}
template <int BITS, int THREADS> constexpr size_t rk_smem_bytes() {
    // packed counters + sorted tile + per-run {start, global slot}
    return (size_t) ((1u << BITS) * (THREADS / 2) + rk_stage_words<BITS, THREADS>() + 2 * (2 * (1u << BITS) + 4)) * 4;
}

struct RkArgs {
    const uint32_t *in0, *in1, *table;
    TileGeom geom;
    uint32_t shift, mask, bins, width, ib, kmax, index_base;
    uint32_t *out0, *out1;
    uint32_t halves; // count table entries per (digit, tile): 2 (6-bit ranking kernel) or 1 (wide digits)
};

/// Staging slot of run r (sorted-tile slot `ls`, first global slot at word address
/// `gword`): 16-byte aligned staging words coincide with 16-byte aligned global
/// addresses, consecutive runs keep RK_PAD words of slack.
B200_DEVICE uint32_t rk_stage_pos(uint32_t ls, uint32_t r, uint32_t gword) {
    return ((ls + RK_PAD * r + 3u) & ~3u) + (gword & 3u);
}

/// in0 / in1: keys (RAW*, PAIRS) or packed words (PACKED) / indices (PAIRS).
/// out0 / out1: permutation (FINAL), keys / indices (PAIRS), packed words (PACKED).
/// ib: index bits of a packed word, width: bits of this digit, index_base: index of
/// element 0 (the permutation holds GLOBAL indices, resources/mkperm.cuh:373-376).
///
/// Counters: the 16-bit counter of (digit d, thread t) lives in 32-bit word
/// d * HALF + (t % HALF), low half for t < HALF = THREADS / 2, high half above -- the
/// counter address is ONE multiply-add away from the digit.  The sorted tile consists
/// of 2 * NB runs (lower-half threads: digits 0 .. NB - 1, then the upper-half
/// threads); the two runs of a digit are adjacent in the OUTPUT, where only the order
/// matters.
///
/// Every run is staged in shared memory so that its 16-byte aligned part coincides
/// with 16-byte aligned global addresses and leaves with ONE bulk copy shared ->
/// global (cp.async.bulk) plus at most 3 + 3 scalar stores for its ragged ends, all
/// issued by ONE thread per run -- there is no per-element store loop.
template <int BITS, int IN, int OUT, int THREADS, bool FULL>
B200_DEVICE void rk_tile(const RkArgs &a, uint32_t tile, uint32_t base, uint32_t tile_count_in,
                         uint32_t ahead_tiles, uint32_t *rk_smem) {
    constexpr uint32_t ITEMS = RK_ITEMS, TILE = THREADS * ITEMS, HALF = THREADS / 2, NWARPS = THREADS / 32;
    constexpr uint32_t NB = 1u << BITS, WPT = NB / 2, G = WPT / 4;
    constexpr uint32_t ROTM = G < HALF / 32 ? G : HALF / 32; // distinct rotations (see below)
    static_assert(G >= 1 && HALF % WPT == 0, "counter rows must consist of whole scan segments");
    static_assert(2 * NB <= THREADS && NWARPS <= 32, "one thread per run");
    uint32_t *s_cnt = rk_smem;                                     // NB * HALF packed counters
    uint32_t *s_stage = rk_smem + NB * HALF;                       // sorted tile (+ alignment slack)
    uint32_t *s_rstart = s_stage + rk_stage_words<BITS, THREADS>();// 2 * NB + 1 run starts (sorted-tile slots)
    uint32_t *s_gpos = s_rstart + 2 * NB + 4;                      // first global slot of the run
    __shared__ uint32_t s_warp[32];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile_count = FULL ? TILE : tile_count_in;
    const uint32_t first = tid * ITEMS; // tile-local index of the thread's first key
    const uint32_t half = tid / HALF;

    #pragma unroll
    for (uint32_t j = 0; j < G; ++j)
        ((uint4 *) s_cnt)[j * THREADS + tid] = make_uint4(0, 0, 0, 0);

    // ---- load (blocked: 64 contiguous bytes per thread)
    uint32_t key[ITEMS];
    auto load = [&](const uint32_t *src, uint32_t (&dst)[ITEMS]) {
        if (FULL && ((uintptr_t) (src + base) & 15) == 0) {
            const uint4 *v = (const uint4 *) (src + base) + tid * (ITEMS / 4);
            #pragma unroll
            for (int q = 0; q < (int) ITEMS / 4; ++q) {
                const uint4 k4 = __ldg(v + q);
                dst[4 * q] = k4.x; dst[4 * q + 1] = k4.y; dst[4 * q + 2] = k4.z; dst[4 * q + 3] = k4.w;
            }
        } else {
            #pragma unroll
            for (int i = 0; i < (int) ITEMS; ++i)
                dst[i] = first + i < tile_count ? __ldg(src + base + first + i) : 0u;
        }
    };
    load(a.in0, key);
    // the next tile of this (persistent) CTA will find its keys in L2
    if (tile + ahead_tiles < a.geom.ntiles && (tid & 1) == 0) {
        uint32_t abase, acount;
        tile_range<TILE>(a.geom, tile + ahead_tiles, abase, acount);
        if (first + 2 * ITEMS <= acount) { // one 128-byte line per two threads
            asm volatile("prefetch.global.L2 [%0];" :: "l"(a.in0 + abase + first));
            if constexpr (IN == RK_PAIRS)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a.in1 + abase + first));
        }
    }
    // first output slot of run tid: digit (tid % NB) of half (tid / NB)
    if (tid < 2 * NB) {
        uint32_t gstart = 0;
        if ((tid & (NB - 1)) < a.bins) {
            const uint32_t group_start = a.geom.tpg == a.geom.ntiles ? 0u : (tile / a.geom.tpg) * a.geom.group_size;
            gstart = group_start + __ldg(a.table + table_index(a.geom, tile, tid & (NB - 1), a.bins) + tid / NB);
        }
        s_gpos[tid] = gstart;
    }

    // The scan below has thread t' rake the WPT consecutive words [t' * WPT, (t' + 1)
    // * WPT) with 128-bit accesses; rotating the 16-byte groups of a thread's segment
    // by rot(t') = (t' * G / 8) % ROTM makes those accesses bank-conflict free (2-way
    // for 6-bit digits).  For the counting accesses the rotation is the same for all
    // lanes of a warp and for every digit, i.e. it is a fixed permutation of the
    // thread's column.
    const uint32_t col = tid & (HALF - 1);
    const uint32_t pcol = (col & ~(WPT - 1)) | (((((col % WPT) >> 2) + (col >> 5) % ROTM) % G) << 2) | (col & 3);
    uint32_t *cnt_mine = s_cnt + pcol;
    const uint32_t inc = 1u << (16 * half);
    if constexpr (IN == RK_RAW1 || IN == RK_RAW) {
        #pragma unroll
        for (int i = 0; i < (int) ITEMS; ++i)
            key[i] = min(key[i], a.kmax);
    }
    auto digit = [&](uint32_t x) -> uint32_t {
        if constexpr (IN == RK_RAW1)
            return x;
        else
            return (x >> a.shift) & a.mask;
    };
    __syncthreads();

    // ---- count: one fire-and-forget atomic per key on the thread's own counters
    #pragma unroll
    for (int i = 0; i < (int) ITEMS; ++i) {
        if (FULL || first + i < tile_count)
            atomicAdd(cnt_mine + digit(key[i]) * HALF, inc);
    }
    __syncthreads();

    // ---- packed exclusive scan of the counters (digit-major, column-minor)
    const uint32_t rot = (tid * G / 8) % ROTM;
    uint4 *seg = (uint4 *) s_cnt + tid * G;
    const uint32_t row = tid * WPT / HALF; // digit of this thread's segment
    uint32_t run;
    {
        uint32_t mine = 0;
        #pragma unroll
        for (uint32_t j = 0; j < G; ++j) {
            const uint4 v = seg[(j + rot) % G];
            mine += v.x + v.y + v.z + v.w;
        }
        uint32_t incl = mine;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= (uint32_t) d)
                incl += up;
        }
        if (lane == 31)
            s_warp[warp] = incl;
        __syncthreads();
        const uint32_t wt = lane < NWARPS ? s_warp[lane] : 0u;
        const uint32_t total = __reduce_add_sync(FULL_MASK, wt);
        run = incl - mine + __reduce_add_sync(FULL_MASK, lane < warp ? wt : 0u);
        run += total << 16; // the upper-half threads' keys come after all lower-half ones
        // run starts: the segment that begins a digit's row of HALF words
        if ((tid * WPT) % HALF == 0) {
            s_rstart[row] = run & 0xffffu;
            s_rstart[NB + row] = run >> 16;
        }
        if (tid == 0)
            s_rstart[2 * NB] = tile_count;
    }
    uint32_t idx[IN == RK_PAIRS ? ITEMS : 1];
    if constexpr (IN == RK_PAIRS)
        load(a.in1, idx);
    // the bulk copies of the CTA's previous tile may still be reading the staging area
    if (tid < 2 * NB)
        bulk_wait_read<0>();
    __syncthreads();

    // ---- second half of the scan: counter = first STAGING slot of its (digit, thread)
    {
        const uint32_t gword = (uint32_t) ((uintptr_t) a.out0 >> 2);
        const uint32_t lo = s_rstart[row], hi = s_rstart[NB + row];
        run += (rk_stage_pos(lo, row, gword + s_gpos[row]) - lo) |
               ((rk_stage_pos(hi, NB + row, gword + s_gpos[NB + row]) - hi) << 16);
        #pragma unroll
        for (uint32_t j = 0; j < G; ++j) {
            uint4 v = seg[(j + rot) % G], o;
            o.x = run; run += v.x;
            o.y = run; run += v.y;
            o.z = run; run += v.z;
            o.w = run; run += v.w;
            seg[(j + rot) % G] = o;
        }
    }
    __syncthreads();

    // ---- place: the atomic returns the key's slot and advances the thread's cursor
    uint32_t slot2[OUT == RK_OUT_PAIRS ? ITEMS / 2 : 1] = {};
    const uint32_t idxmask = a.ib >= 32 ? 0xffffffffu : (1u << a.ib) - 1u;
    const uint32_t hshift = 16 * half;
    // (batches of eight: the atomics of a batch are in flight together, their results
    // are consumed by the stores that follow)
    #pragma unroll
    for (int i0 = 0; i0 < (int) ITEMS; i0 += 8) {
        uint32_t sl[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + j;
            sl[j] = 0;
            if (FULL || first + i < tile_count)
                sl[j] = (atomicAdd(cnt_mine + digit(key[i]) * HALF, inc) >> hshift) & 0xffffu;
        }
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + j;
            if (FULL || first + i < tile_count) {
                const uint32_t w = key[i];
                uint32_t v;
                if constexpr (IN == RK_RAW1) {
                    v = a.index_base + base + first + i;
                } else if constexpr (IN == RK_PACKED) {
                    v = OUT == RK_FINAL ? w & idxmask : ((w >> (a.ib + a.width)) << a.ib) | (w & idxmask);
                } else {
                    const uint32_t ix = IN == RK_PAIRS ? idx[i] : a.index_base + base + first + i;
                    if constexpr (OUT == RK_FINAL)
                        v = ix;
                    else if constexpr (OUT == RK_OUT_PACKED)
                        v = ((w >> (a.shift + a.width)) << a.ib) | ix;
                    else
                        v = w; // keys first, indices in a second round
                }
                s_stage[sl[j]] = v;
                if constexpr (OUT == RK_OUT_PAIRS)
                    slot2[i / 2] |= sl[j] << (16 * (i & 1));
            }
        }
    }

    // thread r writes run r
    auto write_runs = [&](uint32_t *out) {
        fence_proxy_async();
        __syncthreads();
        if (tid < 2 * NB) {
            const uint32_t ls = s_rstart[tid], len = s_rstart[tid + 1] - ls;
            if (len) {
                const uint32_t g = s_gpos[tid];
                // (out0 and out1 are allocations of the same alignment class)
                const uint32_t *src = s_stage + rk_stage_pos(ls, tid, (uint32_t) ((uintptr_t) a.out0 >> 2) + g);
                uint32_t *dst = out + g;
                const uint32_t head = min((4u - ((uint32_t) ((uintptr_t) dst >> 2) & 3u)) & 3u, len);
                const uint32_t mid = (len - head) & ~3u;
                if (mid)
                    bulk_s2g(dst + head, src + head, mid * 4);
                for (uint32_t i = 0; i < head; ++i)
                    dst[i] = src[i];
                for (uint32_t i = head + mid; i < len; ++i)
                    dst[i] = src[i];
            }
            bulk_commit();
        }
    };
    write_runs(a.out0);
    if constexpr (OUT == RK_OUT_PAIRS) {
        // second round: the indices travel through the same staging area
        if (tid < 2 * NB)
            bulk_wait_read<0>();
        __syncthreads();
        #pragma unroll
        for (int i = 0; i < (int) ITEMS; ++i) {
            if (FULL || first + i < tile_count)
                s_stage[(slot2[i / 2] >> (16 * (i & 1))) & 0xffffu] =
                    IN == RK_PAIRS ? idx[i] : a.index_base + base + first + i;
        }
        write_runs(a.out1);
    }
}

template <int THREADS> constexpr int rk_ctas() { return 1024 / THREADS; } // resident CTAs per SM (64 registers)

template <int BITS, int IN, int OUT, int THREADS>
__global__ void __launch_bounds__(THREADS, rk_ctas<THREADS>())
mkperm_rank_place_kernel(const RkArgs a, uint32_t ahead_tiles) {
    constexpr bool TWO = IN == RK_RAW || IN == RK_PAIRS;
    constexpr uint32_t TILE = THREADS * RK_ITEMS;
    static_assert(BITS >= 3 && BITS <= 6, "3..6-bit digits");
    static_assert(TWO ? OUT != RK_FINAL || IN == RK_PAIRS : OUT != RK_OUT_PAIRS, "unsupported combination");
    extern __shared__ __align__(16) uint32_t rk_smem[];
    // persistent CTAs: the bulk stores of a tile drain while the next one is loaded,
    // counted and scanned
    for (uint32_t tile = blockIdx.x; tile < a.geom.ntiles; tile += gridDim.x) {
        uint32_t base, count;
        tile_range<TILE>(a.geom, tile, base, count);
        if (count == TILE)
            rk_tile<BITS, IN, OUT, THREADS, true>(a, tile, base, count, ahead_tiles, rk_smem);
        else if (count) // block-uniform
            rk_tile<BITS, IN, OUT, THREADS, false>(a, tile, base, count, ahead_tiles, rk_smem);
    }
    // the staging area must outlive the bulk copies that read it
    bulk_wait_read<0>();
}

/// Keys per tile: 8192 (512 threads) unless B200_MKPERM_TILE=4096 (development switch)
static uint32_t rk_tile_keys() {
    static int tile = -1;
    if (tile < 0) {
        const char *e = getenv("B200_MKPERM_TILE");
        tile = e && atoi(e) == 4096 ? 4096 : 8192;
    }
    return (uint32_t) tile;
}

template <int BITS, int IN, int OUT, int THREADS>
static cudaError_t rk_launch_cfg(cudaStream_t stream, const RkArgs &a) {
    constexpr size_t smem = rk_smem_bytes<BITS, THREADS>();
    auto kernel = mkperm_rank_place_kernel<BITS, IN, OUT, THREADS>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess)
        return err;
    const uint32_t per_sm = std::min<uint32_t>(rk_ctas<THREADS>(), (uint32_t) ((227 * 1024) / (smem + 1024)));
    const uint32_t grid = std::min<uint32_t>(a.geom.ntiles, per_sm * (uint32_t) sm_count());
    kernel<<<grid, THREADS, smem, stream>>>(a, grid);
    count_launch();
    return cudaGetLastError();
}

template <int BITS, int IN, int OUT>
static cudaError_t rk_launch_one(cudaStream_t stream, const RkArgs &a) {
    if (rk_tile_keys() == 4096)
        return rk_launch_cfg<BITS, IN, OUT, 256>(stream, a);
    return rk_launch_cfg<BITS, IN, OUT, 512>(stream, a);
}

template <int BITS>
static cudaError_t rk_launch_bits(cudaStream_t stream, int in, int out, const RkArgs &a) {
    switch (in * 4 + out) {
        case RK_RAW1 * 4 + RK_FINAL: return rk_launch_one<BITS, RK_RAW1, RK_FINAL>(stream, a);
        case RK_RAW * 4 + RK_OUT_PAIRS: return rk_launch_one<BITS, RK_RAW, RK_OUT_PAIRS>(stream, a);
        case RK_RAW * 4 + RK_OUT_PACKED: return rk_launch_one<BITS, RK_RAW, RK_OUT_PACKED>(stream, a);
        case RK_PAIRS * 4 + RK_FINAL: return rk_launch_one<BITS, RK_PAIRS, RK_FINAL>(stream, a);
        case RK_PAIRS * 4 + RK_OUT_PAIRS: return rk_launch_one<BITS, RK_PAIRS, RK_OUT_PAIRS>(stream, a);
        case RK_PAIRS * 4 + RK_OUT_PACKED: return rk_launch_one<BITS, RK_PAIRS, RK_OUT_PACKED>(stream, a);
        case RK_PACKED * 4 + RK_FINAL: return rk_launch_one<BITS, RK_PACKED, RK_FINAL>(stream, a);
        case RK_PACKED * 4 + RK_OUT_PACKED: return rk_launch_one<BITS, RK_PACKED, RK_OUT_PACKED>(stream, a);
    }
    return cudaErrorInvalidValue;
}

static cudaError_t rk_launch(cudaStream_t stream, int bits, int in, int out, const RkArgs &a) {
    switch (bits) {
        case 3: return rk_launch_bits<3>(stream, in, out, a);
        case 4: return rk_launch_bits<4>(stream, in, out, a);
        case 5: return rk_launch_bits<5>(stream, in, out, a);
        case 6: return rk_launch_bits<6>(stream, in, out, a);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------- single sorting group: ranked tiles, wide digits
//
// 7- and 8-bit digits: thread-private counters (above) would need 256 digits x 512
// threads x 2 bytes.  Here the counters are private to a WARP (16 x 256 words), and
// the order of equal digits inside one warp step -- which a plain atomic counter
// would leave to the hardware -- is made explicit: every lane ORs its lane bit into
// the step's match word of its digit (atoms.or: 5 issue cycles when the lanes
// disagree, against 256 for match.any, tools/ubench_warp_ops.cu), reads the word
// back and ranks itself by the number of lower lanes in it; the first lane of every
// digit adds the group to the warp's counter and clears the word.  A warp walks its
// 512 keys in 16 such steps (step i, lane l <-> key i * 32 + l: tile order), so the
// ranks are stable.  One pass over the [warp][digit] counters (512 threads, 24
// shared-memory accesses each) turns them into the first slot of every (warp, digit)
// pair in the sorted tile; the sorted tile leaves in 16 slices of 512 slots, one per
// warp, the digit of every slot (one byte, written beside the element) selecting the
// run's shift to its global position (coalesced stores within a run; balanced
// whatever the digit distribution).
// Used for 65 .. 256 buckets (one pass); see rw_wanted() for why not beyond.
static constexpr int RW_THREADS = 512, RW_WARPS = RW_THREADS / 32, RW_ITEMS = 16;
static constexpr uint32_t RW_NB = 256, RW_TILE = RW_THREADS * RW_ITEMS, RW_WARP_KEYS = 32 * RW_ITEMS;
static_assert(RW_TILE == 8192, "the tile geometry of the count table");

template <int OUT> constexpr size_t rw_smem_bytes() {
    // match words + counters [warp][digit], sorted tile (two arrays for pairs), run shifts, slot digits
    return (size_t) (2 * RW_WARPS * RW_NB + (OUT == RK_OUT_PAIRS ? 2 : 1) * RW_TILE + RW_NB) * 4 + RW_TILE;
}

template <int IN, int OUT>
__global__ void __launch_bounds__(RW_THREADS, 2)
mkperm_rank_wide_kernel(const RkArgs a) {
    constexpr uint32_t NB = RW_NB, ITEMS = RW_ITEMS, HALFW = RW_WARPS / 2;
    static_assert(IN == RK_RAW || IN == RK_PAIRS ? OUT != RK_FINAL || IN == RK_PAIRS : OUT != RK_OUT_PAIRS,
                  "unsupported combination");
    extern __shared__ __align__(16) uint32_t rw_smem[];
    uint32_t *s_mask = rw_smem;                                              // [warp][digit]
    uint32_t *s_cnt = s_mask + RW_WARPS * NB;                                // [warp][digit]
    uint32_t *s_stage = s_cnt + RW_WARPS * NB;                               // sorted tile
    uint32_t *s_stage1 = s_stage + RW_TILE;                                  // (pairs: the indices)
    uint32_t *s_gdelta = s_stage + (OUT == RK_OUT_PAIRS ? 2 : 1) * RW_TILE;  // global slot - sorted-tile slot
    uint8_t *s_dig = (uint8_t *) (s_gdelta + NB);                            // digit of every sorted slot
    __shared__ uint32_t s_wsum[RW_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bit = 1u << lane, lt = bit - 1u;
    uint32_t *my_mask = s_mask + warp * NB, *my_cnt = s_cnt + warp * NB;
    // the match words start out zero and every step leaves them zero again
    #pragma unroll
    for (uint32_t i = lane; i < NB; i += 32) {
        my_mask[i] = 0;
        my_cnt[i] = 0;
    }
    __syncwarp();
    const uint32_t idxmask = a.ib >= 32 ? 0xffffffffu : (1u << a.ib) - 1u;
    auto digit = [&](uint32_t x) -> uint32_t {
        if constexpr (IN == RK_RAW1)
            return x;
        else
            return (x >> a.shift) & a.mask;
    };

    for (uint32_t tile = blockIdx.x; tile < a.geom.ntiles; tile += gridDim.x) {
        uint32_t base, count;
        tile_range<RW_TILE>(a.geom, tile, base, count);
        if (count == 0) // block-uniform
            continue;
        const bool full = count == RW_TILE;
        const uint32_t first = warp * RW_WARP_KEYS + lane; // tile-local index of this lane's key of step 0

        // ---- load: step i of a warp = 32 consecutive keys
        uint32_t key[ITEMS], idx[IN == RK_PAIRS ? ITEMS : 1];
        #pragma unroll
        for (int i = 0; i < (int) ITEMS; ++i) {
            const uint32_t li = first + i * 32;
            key[i] = full || li < count ? __ldg(a.in0 + base + li) : 0u;
            if constexpr (IN == RK_PAIRS)
                idx[i] = full || li < count ? __ldg(a.in1 + base + li) : 0u;
        }
        // first output slot of this tile's run of digit tid
        uint32_t gstart = 0;
        if (tid < a.bins) {
            const uint32_t group_start = a.geom.tpg == a.geom.ntiles ? 0u : (tile / a.geom.tpg) * a.geom.group_size;
            gstart = group_start + __ldg(a.table + (table_index(a.geom, tile, tid, a.bins) >> 1));
        }
        if constexpr (IN == RK_RAW1 || IN == RK_RAW) {
            #pragma unroll
            for (int i = 0; i < (int) ITEMS; ++i)
                key[i] = min(key[i], a.kmax);
        }

        // ---- rank inside the warp
        uint32_t rk[ITEMS / 2];
        #pragma unroll
        for (int i = 0; i < (int) ITEMS; ++i) {
            const bool valid = full || first + i * 32 < count;
            const uint32_t d = digit(key[i]);
            if (valid)
                atomicOr(my_mask + d, bit);
            __syncwarp();
            const uint32_t m = my_mask[d], c = my_cnt[d];
            const uint32_t r = __popc(m & lt);
            __syncwarp();
            if (valid && r == 0) {
                my_mask[d] = 0;
                my_cnt[d] = c + __popc(m);
            }
            __syncwarp();
            if (i & 1)
                rk[i / 2] |= (c + r) << 16;
            else
                rk[i / 2] = c + r;
        }
        __syncthreads(); // (the previous tile's slices have left, too)

        // ---- counters -> first sorted-tile slot of every (warp, digit) pair.  Thread
        // (dg, sub) owns digit dg of the warps [8 sub, 8 sub + 8).
        {
            const uint32_t dg = tid & (NB - 1), sub = tid / NB;
            static_assert(RW_THREADS == 2 * NB, "two threads per digit");
            uint32_t c8[HALFW], own = 0, other = 0;
            #pragma unroll
            for (uint32_t j = 0; j < HALFW; ++j) {
                c8[j] = s_cnt[(sub * HALFW + j) * NB + dg];
                own += c8[j];
            }
            #pragma unroll
            for (uint32_t j = 0; j < HALFW; ++j)
                other += s_cnt[((sub ^ 1u) * HALFW + j) * NB + dg];
            const uint32_t total = own + other;
            uint32_t incl = total; // scan over the digits (both halves of the CTA, redundantly)
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= (uint32_t) d)
                    incl += up;
            }
            if (lane == 31)
                s_wsum[warp] = incl;
            __syncthreads();
            const uint32_t wt = lane < (warp & (HALFW - 1)) ? s_wsum[sub * HALFW + lane] : 0u;
            const uint32_t dstart = __reduce_add_sync(FULL_MASK, wt) + incl - total;
            uint32_t running = dstart + (sub ? other : 0u);
            #pragma unroll
            for (uint32_t j = 0; j < HALFW; ++j) {
                s_cnt[(sub * HALFW + j) * NB + dg] = running;
                running += c8[j];
            }
            if (sub == 0)
                s_gdelta[dg] = gstart - dstart;
        }
        __syncthreads();

        // ---- place
        #pragma unroll
        for (int i = 0; i < (int) ITEMS; ++i) {
            const uint32_t li = first + i * 32;
            if (full || li < count) {
                const uint32_t w = key[i], d = digit(w);
                const uint32_t slot = my_cnt[d] + ((rk[i / 2] >> (16 * (i & 1))) & 0xffffu);
                s_dig[slot] = (uint8_t) d;
                uint32_t v;
                if constexpr (IN == RK_RAW1) {
                    v = a.index_base + base + li;
                } else if constexpr (IN == RK_PACKED) {
                    v = OUT == RK_FINAL ? w & idxmask : ((w >> (a.ib + a.width)) << a.ib) | (w & idxmask);
                } else {
                    const uint32_t ix = IN == RK_PAIRS ? idx[i] : a.index_base + base + li;
                    if constexpr (OUT == RK_FINAL)
                        v = ix;
                    else if constexpr (OUT == RK_OUT_PACKED)
                        v = ((w >> (a.shift + a.width)) << a.ib) | ix;
                    else {
                        v = w;
                        s_stage1[slot] = ix;
                    }
                }
                s_stage[slot] = v;
            }
        }
        __syncthreads();

        // ---- the warp's counters are its own again
        #pragma unroll
        for (uint32_t i = lane; i < NB; i += 32)
            my_cnt[i] = 0;

        // ---- write: warp w moves the sorted slots [512 w, 512 w + 512): consecutive
        // lanes hold consecutive slots, i.e. (mostly) consecutive addresses of one run
        #pragma unroll 4
        for (int i = 0; i < (int) ITEMS; ++i) {
            const uint32_t k = first + i * 32;
            if (full || k < count) {
                const uint32_t to = s_gdelta[s_dig[k]] + k;
                a.out0[to] = s_stage[k];
                if constexpr (OUT == RK_OUT_PAIRS)
                    a.out1[to] = s_stage1[k];
            }
        }
    }
}

template <int IN, int OUT>
static cudaError_t rw_launch_cfg(cudaStream_t stream, const RkArgs &a) {
    constexpr size_t smem = rw_smem_bytes<OUT>();
    auto kernel = mkperm_rank_wide_kernel<IN, OUT>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess)
        return err;
    const uint32_t grid = std::min<uint32_t>(a.geom.ntiles, 2u * (uint32_t) sm_count());
    kernel<<<grid, RW_THREADS, smem, stream>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t rw_launch(cudaStream_t stream, int in, int out, const RkArgs &a) {
    switch (in * 4 + out) {
        case RK_RAW1 * 4 + RK_FINAL: return rw_launch_cfg<RK_RAW1, RK_FINAL>(stream, a);
        case RK_RAW * 4 + RK_OUT_PAIRS: return rw_launch_cfg<RK_RAW, RK_OUT_PAIRS>(stream, a);
        case RK_RAW * 4 + RK_OUT_PACKED: return rw_launch_cfg<RK_RAW, RK_OUT_PACKED>(stream, a);
        case RK_PAIRS * 4 + RK_FINAL: return rw_launch_cfg<RK_PAIRS, RK_FINAL>(stream, a);
        case RK_PAIRS * 4 + RK_OUT_PAIRS: return rw_launch_cfg<RK_PAIRS, RK_OUT_PAIRS>(stream, a);
        case RK_PAIRS * 4 + RK_OUT_PACKED: return rw_launch_cfg<RK_PAIRS, RK_OUT_PACKED>(stream, a);
        case RK_PACKED * 4 + RK_FINAL: return rw_launch_cfg<RK_PACKED, RK_FINAL>(stream, a);
        case RK_PACKED * 4 + RK_OUT_PACKED: return rw_launch_cfg<RK_PACKED, RK_OUT_PACKED>(stream, a);
    }
    return cudaErrorInvalidValue;
}

/// Digits wider than six bits for this many key bits?  (B200_MKPERM_WIDE: 0 never, 2
/// whenever the key has more than six bits; development switch)
static bool rw_wanted(uint32_t total_bits) {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("B200_MKPERM_WIDE");
        mode = e ? atoi(e) : 1;
    }
    if (mode == 0 || total_bits <= 6 || rk_tile_keys() != RW_TILE)
        return false;
    if (mode == 2)
        return true;
    // 7-8 bits: one wide pass (0.44 ms for 2^26 keys) instead of two 4-bit ones (0.48 ms).
    // Beyond that the wide passes lose: a 16-bit key takes 1.11 ms in two 8-bit passes
    // against 0.94 ms in three of 5 / 5 / 6 bits -- per key and pass the warp-private
    // tables cost 33 shared-memory wavefronts per 32 keys (every access hits a random
    // bank: 3.4-way conflicts on average), the thread-private columns 9 to 11.
    return total_bits <= 8;
}

template <uint32_t TILE>
static void tile_hist_launch(cudaStream_t stream, int form, const uint32_t *in0, const RkArgs &a) {
    const uint32_t grid = (uint32_t) ceil_div(a.geom.ntiles, MT_GROUP / 2);
    const size_t smem = (size_t) MT_GROUP * a.bins * 4;
    uint32_t *table = const_cast<uint32_t *>(a.table);
    switch (form) {
        case RK_RAW1:
            mkperm_tile_hist_kernel<RK_RAW1, TILE><<<grid, MT_THREADS, smem, stream>>>(
                in0, a.geom, a.shift, a.mask, a.kmax, a.bins, a.halves, table);
            break;
        case RK_RAW:
            mkperm_tile_hist_kernel<RK_RAW, TILE><<<grid, MT_THREADS, smem, stream>>>(
                in0, a.geom, a.shift, a.mask, a.kmax, a.bins, a.halves, table);
            break;
        default:
            mkperm_tile_hist_kernel<RK_PACKED, TILE><<<grid, MT_THREADS, smem, stream>>>(
                in0, a.geom, a.shift, a.mask, a.kmax, a.bins, a.halves, table);
    }
    count_launch();
}

static int histogram_launch(cudaStream_t stream, const uint32_t *keys, uint64_t size,
                            uint32_t bins, uint32_t *hist) {
    B200_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t) bins * sizeof(uint32_t), stream));
    const int sms = sm_count();
    constexpr uint32_t SWEEP_BINS = 49152; // 192 KiB of counters per CTA
    if (bins <= 8 * SWEEP_BINS) {
        // keys beyond one table: one sweep over the array per sub-range of the keys
        const uint32_t sweeps = (uint32_t) ceil_div(bins, SWEEP_BINS);
        const uint32_t per = (uint32_t) ceil_div(bins, sweeps);
        const size_t smem = (size_t) per * sizeof(uint32_t);
        if (smem > 48 * 1024)
            B200_CUDA_CHECK(cudaFuncSetAttribute(histogram_smem_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int) smem));
        uint32_t per_sm = smem <= 32 * 1024 ? 2 : 1;
        uint32_t grid = (uint32_t) std::max<uint64_t>(
            1, std::min<uint64_t>((uint64_t) sms * per_sm, ceil_div(size, 4096)));
        for (uint32_t lo = 0; lo < bins; lo += per) {
            histogram_smem_kernel<<<grid, 1024, smem, stream>>>(keys, size, lo, std::min(per, bins - lo),
                                                                bins, hist);
            count_launch();
        }
    } else {
        uint32_t grid = (uint32_t) std::max<uint64_t>(
            1, std::min<uint64_t>((uint64_t) sms * 8, ceil_div(size, 256)));
        histogram_global_kernel<<<grid, 256, 0, stream>>>(keys, size, bins, hist);
        count_launch();
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return cuda_fail(err, "kernel launch");
    return B200_OK;
}

/// Bucket-count histogram: one shared-memory table per CTA when it fits, a few
/// sweeps over key sub-ranges when it does not (global atomics only beyond that).
static int histogram_launch(cudaStream_t stream, const uint32_t *keys, uint64_t size,
                            uint32_t bins, uint32_t *hist);

/// block_size == size: ranked tile passes (see mkperm_rank_place_kernel)
/// Ranked-tile sort of every group of 'group_size' consecutive keys (group_size ==
/// size: one group, the vcall case).  offsets: single group only.
static int mkperm_ranked(cudaStream_t stream, const uint32_t *values, uint32_t size, uint32_t group_size,
                         uint32_t bucket_count, uint32_t *perm, uint32_t *offsets, bool by_size) {
    const uint32_t index_base = 0;
    uint32_t total_bits = 1;
    while (total_bits < 32 && (1ull << total_bits) < bucket_count)
        total_bits++;
    uint32_t ib = 1;
    while (ib < 32 && (1ull << ib) < (uint64_t) index_base + size)
        ib++;
    const bool wide = rw_wanted(total_bits);
    const uint32_t npasses = wide ? (total_bits + 7) / 8 : (total_bits + 5) / 6;
    const uint32_t halves = wide ? 1 : 2;
    TileGeom geom{};
    geom.size = size;
    geom.group_size = group_size;
    const uint32_t tile_keys = rk_tile_keys();
    geom.tpg = (uint32_t) ceil_div(group_size, tile_keys);
    const uint64_t ngroups = ceil_div(size, group_size);
    if (ngroups * geom.tpg > 0x7fffffffull)
        return fail(B200_ERR_INVALID, "jit_block_mkperm(): too many tiles!");
    geom.ntiles = (uint32_t) (ngroups * geom.tpg);
    const uint32_t ntiles = geom.ntiles;

    // digit widths: as even as possible, wider digits last (a wide digit costs
    // counters and runs; the last passes move one word per element, not two)
    uint32_t width[8], shift_of[8];
    for (uint32_t p = 0, sh = 0; p < npasses; ++p) {
        width[p] = total_bits / npasses + (p >= npasses - total_bits % npasses ? 1 : 0);
        shift_of[p] = sh;
        sh += width[p];
    }

    const uint64_t max_counts = (uint64_t) (wide ? 256 : 128) * ntiles;
    uint32_t *table = (uint32_t *) temp_alloc(max_counts * 4, stream);
    uint32_t *tmp[4] = { nullptr, nullptr, nullptr, nullptr };
    auto cleanup = [&]() {
        temp_free(table, stream);
        for (auto t : tmp)
            temp_free(t, stream);
    };
    if (!table)
        return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");

    int form = npasses == 1 ? RK_RAW1 : RK_RAW; // how elements reach the current pass
    const uint32_t *in0 = values, *in1 = nullptr;
    for (uint32_t p = 0; p < npasses; ++p) {
        const bool last = p + 1 == npasses;
        const uint32_t consumed = shift_of[p] + width[p];
        RkArgs a{};
        a.in0 = in0;
        a.in1 = in1;
        a.table = table;
        a.geom = geom;
        a.ib = ib;
        a.kmax = bucket_count - 1;
        a.width = width[p];
        a.index_base = index_base;
        a.halves = halves;
        if (form == RK_RAW1) {
            a.shift = 0;
            a.mask = 0xffffffffu;
            a.bins = bucket_count;
        } else {
            a.shift = form == RK_PACKED ? ib : shift_of[p];
            a.mask = (1u << width[p]) - 1u;
            a.bins = last ? ((bucket_count - 1) >> shift_of[p]) + 1 : 1u << width[p];
            a.bins = std::min(a.bins, 1u << width[p]);
        }
        int out;
        if (last)
            out = RK_FINAL;
        else
            out = (total_bits - consumed) + ib <= 32 ? RK_OUT_PACKED : RK_OUT_PAIRS;
        if (last) {
            a.out0 = perm;
        } else {
            const int set = (p & 1) * 2;
            for (int i = set; i < set + (out == RK_OUT_PAIRS ? 2 : 1); ++i) {
                if (!tmp[i])
                    tmp[i] = (uint32_t *) temp_alloc((size_t) size * 4, stream);
                if (!tmp[i]) {
                    cleanup();
                    return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
                }
            }
            a.out0 = tmp[set];
            a.out1 = tmp[set + 1];
        }

        const uint64_t ncounts = (uint64_t) a.bins * halves * ntiles;
        if (tile_keys == 4096)
            tile_hist_launch<4096>(stream, form, in0, a);
        else
            tile_hist_launch<8192>(stream, form, in0, a);
        // one exclusive scan per group over its [digit][tile][half] counts
        int rc = b200_block_prefix_reduce(stream, B200_VT_UINT32, B200_OP_ADD, ncounts,
                                          (uint64_t) a.bins * geom.tpg * halves, 1, 0, table, table);
        if (rc) {
            cleanup();
            return rc;
        }
        // single pass: bucket starts are a strided view of the scanned table
        if (npasses == 1 && offsets && ngroups == 1) {
            uint32_t *records = (uint32_t *) temp_alloc(((size_t) bucket_count * 4 + 1) * 4, stream);
            if (!records) {
                cleanup();
                return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
            }
            mkperm_offsets_kernel<<<1, 1024, 0, stream>>>(table, halves * (uint64_t) ntiles, bucket_count, size, records);
            count_launch();
            rc = deliver_records(stream, records, bucket_count, offsets, by_size);
            temp_free(records, stream);
            if (rc) {
                cleanup();
                return rc;
            }
        }
        const uint32_t bits = form == RK_RAW1 ? std::max(3u, total_bits) : std::max(3u, width[p]);
        cudaError_t err = wide ? rw_launch(stream, form, out, a) : rk_launch(stream, (int) bits, form, out, a);
        if (err != cudaSuccess) {
            cleanup();
            return cuda_fail(err, "mkperm_rank_place_kernel");
        }
        in0 = a.out0;
        in1 = out == RK_OUT_PAIRS ? a.out1 : nullptr;
        form = out == RK_OUT_PAIRS ? RK_PAIRS : RK_PACKED;
    }

    // several passes: the bucket starts are read off the finished permutation
    if (npasses > 1 && offsets && ngroups == 1) {
        size_t hist_words = ((size_t) bucket_count + 3) & ~(size_t) 3; // keep records 16-byte aligned
        const size_t rec_words = ((size_t) bucket_count * 4 + 1 + 3) & ~(size_t) 3;
        uint32_t *hist = (uint32_t *) temp_alloc((hist_words + rec_words + ceil_div(bucket_count, 1024)) * 4, stream);
        if (!hist) {
            cleanup();
            return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
        }
        uint32_t *records = hist + hist_words, *block_counts = records + rec_words;
        // bucket starts: 17-ary search over the finished permutation (latency bound,
        // independent of the bucket count) instead of a histogram of the keys (49 us + a
        // scan for <= 49152 buckets, a sweep per 49152 buckets beyond)
        int rc = B200_OK;
        mkperm_bounds_kernel<<<(uint32_t) ceil_div((uint64_t) bucket_count * 16, 256), 256, 0, stream>>>(
            values, perm, size, index_base, bucket_count, hist);
        count_launch();
        if (!rc) {
            offsets_from_starts(stream, hist, bucket_count, size, records, block_counts);
            rc = deliver_records(stream, records, bucket_count, offsets, by_size);
        }
        temp_free(hist, stream);
        if (rc) {
            cleanup();
            return rc;
        }
    }
    cleanup();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return cuda_fail(err, "kernel launch");
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_mkperm_histogram(void *stream_, const uint32_t *values, uint64_t size,
                          uint32_t bucket_count, uint32_t *hist) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (bucket_count == 0)
        return fail(B200_ERR_INVALID, "jit_block_mkperm(): bucket_count cannot be zero!");
    cudaStream_t stream = resolve_stream(stream_);
    if (size == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t) bucket_count * 4, stream));
        return B200_OK;
    }
    return histogram_launch(stream, values, size, bucket_count, hist);
}

static int mkperm_impl(void *stream_, const uint32_t *values, uint32_t size,
                       uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                       uint32_t *offsets, uint32_t *unique, bool wait, bool by_size = false);

int b200_block_mkperm(void *stream_, const uint32_t *values, uint32_t size,
                      uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                      uint32_t *offsets, uint32_t *unique) {
    return mkperm_impl(stream_, values, size, block_size, bucket_count, perm, offsets, unique, true);
}

int b200_block_mkperm_async(void *stream_, const uint32_t *values, uint32_t size,
                            uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                            uint32_t *offsets) {
    return mkperm_impl(stream_, values, size, block_size, bucket_count, perm, offsets, nullptr, false);
}

int b200_call_reduce(void *stream_, const uint32_t *ids, uint32_t size, uint32_t id_bound,
                     uint32_t *perm, uint32_t *offsets, uint32_t *unique) {
    if (id_bound == 0xffffffffu)
        return fail(B200_ERR_INVALID, "jit_var_call_reduce(): too many callables!");
    if (!offsets)
        return fail(B200_ERR_INVALID, "jit_var_call_reduce(): the bucket records are required!");
    return mkperm_impl(stream_, ids, size, size, id_bound + 1, perm, offsets, unique, true, true);
}

int b200_call_reduce_async(void *stream_, const uint32_t *ids, uint32_t size, uint32_t id_bound,
                           uint32_t *perm, uint32_t *offsets) {
    if (id_bound == 0xffffffffu)
        return fail(B200_ERR_INVALID, "jit_var_call_reduce(): too many callables!");
    if (!offsets)
        return fail(B200_ERR_INVALID, "jit_var_call_reduce(): the bucket records are required!");
    return mkperm_impl(stream_, ids, size, size, id_bound + 1, perm, offsets, nullptr, false, true);
}

static int mkperm_impl(void *stream_, const uint32_t *values, uint32_t size,
                       uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                       uint32_t *offsets, uint32_t *unique, bool wait, bool by_size) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (unique)
        *unique = 0;
    // src/cuda_ts.cpp:792-795
    if (size == 0)
        return B200_OK;
    if (bucket_count == 0)
        return fail(B200_ERR_INVALID, "jit_block_mkperm(): bucket_count cannot be zero!");
    if (block_size == 0)
        return fail(B200_ERR_INVALID, "jit_block_mkperm(): block_size cannot be zero!");
    if (block_size > size)
        block_size = size;

    const uint64_t ngroups = ceil_div(size, block_size);
    // NOTE JitFlag::ForbidSynchronization: with offsets the call waits, but the
    // reference waits on an EVENT (cuEventSynchronize, src/cuda_ts.cpp:964-967), not
    // through jitc_sync_thread, so it does not raise under that flag -- neither does
    // this entry point (jit_compress, jit_all / jit_any and jit_memcpy do).
    cudaStream_t stream = resolve_stream(stream_);
    const int sms = sm_count();
    HistoryScope hs(stream, B200_KERNEL_MKPERM, size);
    NvtxRange nvtx("jit_block_mkperm"); // ProfilerPhase of src/util.cpp:218-229

    // one sorting group (vectorised method dispatch) or groups of at least half a tile:
    // ranked tiles.  Smaller groups would leave the tiles mostly empty: row kernels.
    if (ngroups == 1 || block_size >= rk_tile_keys() / 2) {
        rc = mkperm_ranked(stream, values, size, block_size, bucket_count, perm,
                           ngroups == 1 ? offsets : nullptr, by_size);
        if (rc)
            return rc;
        if (wait && offsets && ngroups == 1) {
            // the reference waits on an event here (src/cuda_ts.cpp:964-967)
            B200_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (unique)
                *unique = offsets[4 * (size_t) bucket_count];
        }
        return B200_OK;
    }

    // ---- digit plan
    uint32_t total_bits = 0;
    while (total_bits < 32 && (1ull << total_bits) < bucket_count)
        total_bits++;
    const uint32_t digit_bits = 11;
    uint32_t npasses = 1, bits_per = 0;
    if (bucket_count > MKPERM_MAX_BINS) {
        npasses = (total_bits + digit_bits - 1) / digit_bits;
        bits_per = (total_bits + npasses - 1) / npasses;
    }

    // ---- row geometry (shared by all passes)
    uint32_t max_bins = npasses == 1 ? bucket_count : (1u << bits_per);
    uint64_t target_rows = (uint64_t) sms * (max_bins <= 256 ? 32 : 16);
    uint64_t rows_per_group = 1, row_len = block_size;
    if (ngroups < target_rows) {
        uint64_t want = ceil_div(target_rows, ngroups);
        row_len = ceil_div(ceil_div(block_size, want), 256) * 256;
        rows_per_group = ceil_div(block_size, row_len);
    }

    size_t smem = (size_t) MKPERM_WARPS * max_bins * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_hist_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_place_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }

    // groups per launch so that the count table stays below 1 GiB
    uint64_t counts_per_group = (uint64_t) max_bins * rows_per_group;
    uint64_t groups_per_launch = std::max<uint64_t>(1, (1ull << 28) / counts_per_group);
    groups_per_launch = std::min(groups_per_launch, ngroups);
    if (counts_per_group >= (1ull << 31))
        return fail(B200_ERR_INVALID, "jit_block_mkperm(): bucket_count too large!");

    uint32_t *counts = (uint32_t *) temp_alloc(groups_per_launch * counts_per_group * 4, stream);
    uint32_t *tmp[4] = { nullptr, nullptr, nullptr, nullptr }; // keysA idxA keysB idxB
    bool ok = counts != nullptr;
    if (npasses > 1) {
        for (int i = 0; i < (npasses > 2 ? 4 : 2) && ok; ++i) {
            tmp[i] = (uint32_t *) temp_alloc((size_t) size * 4, stream);
            ok &= tmp[i] != nullptr;
        }
    }
    auto cleanup = [&]() {
        temp_free(counts, stream);
        for (auto t : tmp)
            temp_free(t, stream);
    };
    if (!ok) {
        cleanup();
        return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
    }

    for (uint32_t pass = 0; pass < npasses; ++pass) {
        MkpermGeom g{};
        g.size = size;
        g.group_size = block_size;
        g.rows_per_group = (uint32_t) rows_per_group;
        g.row_len = row_len;
        if (npasses == 1) {
            g.bins = bucket_count;
            g.shift = 0;
            g.mask = 0xffffffffu;
        } else {
            g.shift = pass * bits_per;
            g.mask = (1u << bits_per) - 1;
            uint64_t remaining = ceil_div(bucket_count, 1ull << g.shift);
            g.bins = (uint32_t) std::min<uint64_t>(1ull << bits_per, remaining);
        }
        const bool last = pass + 1 == npasses;
        const uint32_t *keys_in = pass == 0 ? values : tmp[((pass - 1) & 1) * 2];
        const uint32_t *idx_in = pass == 0 ? nullptr : tmp[((pass - 1) & 1) * 2 + 1];
        uint32_t *keys_out = last ? nullptr : tmp[(pass & 1) * 2];
        uint32_t *idx_out = last ? perm : tmp[(pass & 1) * 2 + 1];
        size_t pass_smem = (size_t) MKPERM_WARPS * g.bins * sizeof(uint32_t);

        for (uint64_t g0 = 0; g0 < ngroups; g0 += groups_per_launch) {
            g.group0 = g0;
            g.ngroups = (uint32_t) std::min(groups_per_launch, ngroups - g0);
            uint64_t nrows = (uint64_t) g.ngroups * rows_per_group;
            uint32_t grid = (uint32_t) ceil_div(nrows, MKPERM_WARPS);
            uint64_t ncounts = (uint64_t) g.ngroups * g.bins * rows_per_group;

            mkperm_hist_kernel<<<grid, MKPERM_THREADS, pass_smem, stream>>>(keys_in, counts, g);
            count_launch();
            rc = b200_block_prefix_reduce(stream, B200_VT_UINT32, B200_OP_ADD, ncounts,
                                          (uint64_t) g.bins * rows_per_group, 1, 0, counts, counts);
            if (rc) {
                cleanup();
                return rc;
            }

            // single group, single pass: bucket starts are a strided view of the scan
            if (last && npasses == 1 && offsets && ngroups == 1) {
                uint32_t *records = (uint32_t *) temp_alloc(((size_t) bucket_count * 4 + 1) * 4, stream);
                if (!records) {
                    cleanup();
                    return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
                }
                mkperm_offsets_kernel<<<1, 1024, 0, stream>>>(counts, rows_per_group, bucket_count,
                                                             size, records);
                count_launch();
                rc = deliver_records(stream, records, bucket_count, offsets, by_size);
                temp_free(records, stream);
                if (rc) {
                    cleanup();
                    return rc;
                }
            }

            mkperm_place_kernel<<<grid, MKPERM_THREADS, pass_smem, stream>>>(
                keys_in, idx_in, counts, idx_out, keys_out, g);
            count_launch();
        }
    }

    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        cleanup();
        return cuda_fail(err, "kernel launch");
    }

    // multi-pass: bucket sizes come from a whole-array histogram of the full key
    if (npasses > 1 && offsets && ngroups == 1) {
        size_t hist_words = ((size_t) bucket_count + 3) & ~(size_t) 3; // keep records 16-byte aligned
        uint32_t *hist = (uint32_t *) temp_alloc((hist_words + (size_t) bucket_count * 4 + 1) * 4, stream);
        if (!hist) {
            cleanup();
            return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
        }
        uint32_t *records = hist + hist_words;
        rc = histogram_launch(stream, values, size, bucket_count, hist);
        if (!rc)
            rc = b200_block_prefix_reduce(stream, B200_VT_UINT32, B200_OP_ADD, bucket_count,
                                          bucket_count, 1, 0, hist, hist);
        if (!rc) {
            mkperm_offsets_kernel<<<1, 1024, 0, stream>>>(hist, 1, bucket_count, size, records);
            count_launch();
            rc = deliver_records(stream, records, bucket_count, offsets, by_size);
        }
        temp_free(hist, stream);
        if (rc) {
            cleanup();
            return rc;
        }
    }

    cleanup();

    if (wait && offsets && ngroups == 1) {
        // the reference waits on an event here (src/cuda_ts.cpp:964-967)
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (unique)
            *unique = offsets[4 * (size_t) bucket_count];
    }
    return B200_OK;
}

} // extern "C"
