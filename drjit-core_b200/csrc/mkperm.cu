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

namespace b200 {

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

/// Whole-array histogram with a per-CTA shared-memory table that is flushed
/// with global atomics (bins * 4 bytes must fit the dynamic shared memory).
__global__ void __launch_bounds__(1024)
histogram_smem_kernel(const uint32_t *__restrict__ keys, uint64_t size, uint32_t bins,
                      uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t mk_smem[];
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x)
        mk_smem[b] = 0;
    __syncthreads();
    uint64_t chunk = (size + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 3) & ~3ull;
    uint64_t start = min((uint64_t) blockIdx.x * chunk, size), end = min(start + chunk, size);
    bool aligned = ((uintptr_t) keys & 15) == 0;
    if (aligned) {
        uint64_t nvec = (end - start) / 4;
        const uint4 *v = (const uint4 *) (keys + start);
        for (uint64_t q = threadIdx.x; q < nvec; q += blockDim.x) {
            uint4 k4 = ld_stream(v + q);
            atomicAdd(&mk_smem[min(k4.x, bins - 1)], 1u);
            atomicAdd(&mk_smem[min(k4.y, bins - 1)], 1u);
            atomicAdd(&mk_smem[min(k4.z, bins - 1)], 1u);
            atomicAdd(&mk_smem[min(k4.w, bins - 1)], 1u);
        }
        start += nvec * 4;
    }
    for (uint64_t k = start + threadIdx.x; k < end; k += blockDim.x)
        atomicAdd(&mk_smem[min(keys[k], bins - 1)], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) {
        uint32_t c = mk_smem[b];
        if (c)
            atomicAdd(&hist[b], c);
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
/// for 4 * bins + 1 words, the last one receiving the number of records.
__global__ void __launch_bounds__(1024)
mkperm_offsets_kernel(const uint32_t *__restrict__ starts, uint64_t stride, uint32_t bins,
                      uint32_t size, uint32_t *__restrict__ records) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_running;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < bins; base += 1024) {
        uint32_t b = base + tid;
        uint32_t st = 0, sz = 0;
        if (b < bins) {
            st = starts[(uint64_t) b * stride];
            uint32_t nxt = b + 1 < bins ? starts[(uint64_t) (b + 1) * stride] : size;
            sz = nxt - st;
        }
        bool flag = sz > 0;
        uint32_t ballot = __ballot_sync(FULL_MASK, flag);
        if (lane == 0)
            warp_sums[warp] = __popc(ballot);
        __syncthreads();
        uint32_t running = s_running;
        uint32_t wsum = warp_sums[lane];
        uint32_t wincl = wsum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t up = __shfl_up_sync(FULL_MASK, wincl, d);
            if (lane >= (uint32_t) d)
                wincl += up;
        }
        uint32_t warp_excl = __shfl_sync(FULL_MASK, wincl - wsum, warp);
        uint32_t total = __shfl_sync(FULL_MASK, wincl, 31);
        if (flag) {
            uint32_t rec = running + warp_excl + __popc(ballot & ((1u << lane) - 1));
            ((uint4 *) records)[rec] = make_uint4(b, st, sz, 0);
        }
        __syncthreads();
        if (tid == 0)
            s_running = running + total;
        __syncthreads();
    }
    if (tid == 0)
        records[4 * (size_t) bins] = s_running;
}

// ------------------------------------------------ single sorting group: tiles
//
// block_size == size (what vectorised method dispatch passes): the array is cut
// into tiles of 8192 keys, one CTA per tile.
//   tile_hist   per-tile digit counts, written digit-major (table[d][tile]) so
//               that ONE exclusive scan of the flattened table (scan_fast.cu)
//               yields the first output slot of every (digit, tile) pair;
//   tile_place  re-reads the tile, ranks its keys stably (a warp walks its 512
//               keys in index order, 32 per step: match.any groups equal digits,
//               the lowest lane advances the warp's private counter), sorts the
//               tile by digit in shared memory and writes every digit's run with
//               contiguous stores.  The per-row variant above has every warp
//               trickle single elements into `bins` output streams at once,
//               which thrashes L2 (measured on B200: 6.5x DRAM write
//               amplification, 2-4 ms for 2^26 keys).
static constexpr int MT_THREADS = 512;
static constexpr int MT_WARPS = MT_THREADS / 32;
static constexpr int MT_ITEMS = 16;
static constexpr uint32_t MT_TILE = MT_THREADS * MT_ITEMS; // 8192 keys
static constexpr uint32_t MT_MAX_BINS = 1024;              // 10-bit digits

B200_DEVICE uint32_t mt_digit(uint32_t key, uint32_t shift, uint32_t mask, uint32_t bins) {
    return min((key >> shift) & mask, bins - 1); // out-of-range keys must not corrupt memory
}

/// table[d * ntiles + tile] = number of keys of the tile whose digit is d.  A CTA
/// counts MT_GROUP consecutive tiles into one shared-memory histogram each, so
/// that the MT_GROUP entries of a digit are contiguous in the digit-major table
/// (one 32-byte run instead of MT_GROUP scattered words).
/// Dynamic shared memory: MT_GROUP * bins counters.
static constexpr uint32_t MT_GROUP = 8;

__global__ void __launch_bounds__(MT_THREADS)
mkperm_tile_hist_kernel(const uint32_t *__restrict__ keys, uint64_t size, uint32_t ntiles,
                        uint32_t shift, uint32_t mask, uint32_t bins,
                        uint32_t *__restrict__ table) {
    extern __shared__ uint32_t mk_smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t tile0 = blockIdx.x * MT_GROUP;
    for (uint32_t i = tid; i < MT_GROUP * bins; i += MT_THREADS)
        mk_smem[i] = 0;
    __syncthreads();
    const bool aligned = ((uintptr_t) keys & 15) == 0;
    #pragma unroll 1
    for (uint32_t t = 0; t < MT_GROUP; t += 2) {
        // two tiles per step: 8 x 16-byte loads in flight per thread
        uint4 k4[2][MT_ITEMS / 4];
        bool full[2];
        #pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t base = (uint64_t) (tile0 + t + u) * MT_TILE;
            full[u] = aligned && base + MT_TILE <= size;
            if (full[u]) {
                const uint4 *v = (const uint4 *) (keys + base);
                #pragma unroll
                for (int j = 0; j < MT_ITEMS / 4; ++j)
                    k4[u][j] = ld_stream(v + j * MT_THREADS + tid);
            }
        }
        #pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t *h = mk_smem + (t + u) * bins;
            const uint64_t base = (uint64_t) (tile0 + t + u) * MT_TILE;
            if (full[u]) {
                #pragma unroll
                for (int j = 0; j < MT_ITEMS / 4; ++j) {
                    atomicAdd(&h[mt_digit(k4[u][j].x, shift, mask, bins)], 1u);
                    atomicAdd(&h[mt_digit(k4[u][j].y, shift, mask, bins)], 1u);
                    atomicAdd(&h[mt_digit(k4[u][j].z, shift, mask, bins)], 1u);
                    atomicAdd(&h[mt_digit(k4[u][j].w, shift, mask, bins)], 1u);
                }
            } else if (base < size) {
                #pragma unroll 4
                for (int j = 0; j < MT_ITEMS; ++j) {
                    const uint64_t g = base + (uint64_t) j * MT_THREADS + tid;
                    if (g < size)
                        atomicAdd(&h[mt_digit(__ldg(keys + g), shift, mask, bins)], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < MT_GROUP * bins; i += MT_THREADS) {
        const uint32_t d = i / MT_GROUP, t = i % MT_GROUP;
        if (tile0 + t < ntiles)
            table[(uint64_t) d * ntiles + tile0 + t] = mk_smem[t * bins + d];
    }
}

/// 'table' holds the exclusively scanned counts.  The element moved is
/// idx_in[g] (PAIRS) or the key's own index g; keys_out (optional) receives the
/// key at the same slot for the next digit pass.
/// Dynamic shared memory: 2 * MT_TILE + bins words, then MT_WARPS * (bins + 2)
/// 16-bit counters.
template <bool PAIRS>
__global__ void __launch_bounds__(MT_THREADS, 2)
mkperm_tile_place_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ idx_in,
                         const uint32_t *__restrict__ table, uint64_t size, uint32_t ntiles,
                         uint32_t shift, uint32_t mask, uint32_t bins,
                         uint32_t *__restrict__ perm_out, uint32_t *__restrict__ keys_out) {
    extern __shared__ uint32_t mk_smem[];
    uint32_t *s_key = mk_smem;
    uint32_t *s_idx = mk_smem + MT_TILE;
    uint32_t *s_delta = mk_smem + 2 * MT_TILE;
    uint16_t *s_hist = (uint16_t *) (mk_smem + 2 * MT_TILE + bins);
    // 16-bit counters, two per word; + sentinel bin for the lanes past the end
    const uint32_t hstride = (bins + 3) & ~1u;
    __shared__ uint32_t s_warp[MT_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t tile = blockIdx.x;
    const uint64_t base = (uint64_t) tile * MT_TILE;

    for (uint32_t i = tid; i < MT_WARPS * hstride / 2; i += MT_THREADS)
        ((uint32_t *) s_hist)[i] = 0;

    // ---- load: warp w owns keys [w * 512, (w + 1) * 512) of the tile, striped
    uint32_t key[MT_ITEMS];
    #pragma unroll
    for (int i = 0; i < MT_ITEMS; ++i) {
        const uint64_t g = base + warp * 512 + i * 32 + lane;
        key[i] = g < size ? __ldg(keys + g) : 0u;
    }
    // first output slot of this thread's digits in this tile (used much later)
    const uint32_t dpt = (bins + MT_THREADS - 1) / MT_THREADS; // digits per thread (<= 2)
    const uint32_t d0 = tid * dpt;
    uint32_t gstart[2] = { 0, 0 };
    #pragma unroll
    for (uint32_t q = 0; q < 2; ++q)
        if (q < dpt && d0 + q < bins)
            gstart[q] = __ldg(table + (uint64_t) (d0 + q) * ntiles + tile);
    __syncthreads();

    // ---- stable rank of every key among the warp's keys with the same digit:
    // match.any groups the lanes with equal digits, the lowest lane of a group
    // advances the warp's private counter, everyone takes counter + (number of
    // group members below it).  (A variant that batches the 16 match / atomic /
    // shuffle steps to break the dependency chain measured 20 % slower: more
    // live registers than the 64 that two CTAs per SM allow.)
    uint32_t rank2[MT_ITEMS / 2]; // two 16-bit ranks per register (a rank is < 512)
    #pragma unroll
    for (int i = 0; i < MT_ITEMS / 2; ++i)
        rank2[i] = 0;
    {
        uint16_t *whr = s_hist + warp * hstride;
        #pragma unroll
        for (int i = 0; i < MT_ITEMS; ++i) {
            const uint64_t g = base + warp * 512 + i * 32 + lane;
            const uint32_t d = g < size ? mt_digit(key[i], shift, mask, bins) : bins;
            const uint32_t peers = __match_any_sync(FULL_MASK, d);
            const uint32_t below = __popc(peers & lt_mask);
            uint32_t old = 0;
            if (below == 0) {
                old = whr[d];
                whr[d] = (uint16_t) (old + __popc(peers));
            }
            old = __shfl_sync(FULL_MASK, old, __ffs(peers) - 1);
            rank2[i / 2] |= (old + below) << (16 * (i & 1));
        }
    }
    uint16_t *wh = s_hist + warp * hstride;
    __syncthreads();

    // ---- per digit: total over the warps, exclusive prefix over the digits,
    // then every warp's first slot of the digit in the sorted tile
    uint32_t cnt0 = 0, cnt1 = 0;
    if (d0 < bins) {
        #pragma unroll
        for (int w = 0; w < MT_WARPS; ++w)
            cnt0 += s_hist[w * hstride + d0];
    }
    if (dpt > 1 && d0 + 1 < bins) {
        #pragma unroll
        for (int w = 0; w < MT_WARPS; ++w)
            cnt1 += s_hist[w * hstride + d0 + 1];
    }
    const uint32_t mine = cnt0 + cnt1;
    uint32_t incl = mine;
    #pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL_MASK, incl, s);
        if (lane >= (uint32_t) s)
            incl += up;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    uint32_t lstart = incl - mine; // first slot of digit d0 in the sorted tile
    #pragma unroll
    for (int w = 0; w < MT_WARPS; ++w)
        lstart += (uint32_t) w < warp ? s_warp[w] : 0u;
    #pragma unroll
    for (uint32_t q = 0; q < 2; ++q) {
        const uint32_t d = d0 + q;
        if (q < dpt && d < bins) {
            // sorted-tile slot j of this digit goes to global slot j + delta
            s_delta[d] = gstart[q] - lstart;
            uint32_t run = lstart;
            #pragma unroll 4
            for (int w = 0; w < MT_WARPS; ++w) {
                const uint32_t t = s_hist[w * hstride + d];
                s_hist[w * hstride + d] = (uint16_t) run;
                run += t;
            }
            lstart = run;
        }
    }
    __syncthreads();

    // ---- sort the tile by digit in shared memory
    #pragma unroll
    for (int i = 0; i < MT_ITEMS; ++i) {
        const uint64_t g = base + warp * 512 + i * 32 + lane;
        if (g < size) {
            const uint32_t d = mt_digit(key[i], shift, mask, bins);
            const uint32_t slot = wh[d] + ((rank2[i / 2] >> (16 * (i & 1))) & 0xffffu);
            s_key[slot] = key[i];
            s_idx[slot] = PAIRS ? __ldg(idx_in + g) : (uint32_t) g;
        }
    }
    __syncthreads();

    // ---- write every digit's run with contiguous stores
    const uint32_t tile_count = (uint32_t) min((uint64_t) MT_TILE, size - base);
    for (uint32_t j = tid; j < tile_count; j += MT_THREADS) {
        const uint32_t k = s_key[j];
        const uint32_t pos = j + s_delta[mt_digit(k, shift, mask, bins)];
        perm_out[pos] = s_idx[j];
        if (keys_out)
            keys_out[pos] = k;
    }
}

static size_t tile_place_smem(uint32_t bins) {
    return (size_t) (2 * MT_TILE + bins) * 4 + (size_t) MT_WARPS * ((bins + 3) & ~1u) * 2;
}

static int histogram_launch(cudaStream_t stream, const uint32_t *keys, uint64_t size,
                            uint32_t bins, uint32_t *hist) {
    B200_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t) bins * sizeof(uint32_t), stream));
    const int sms = sm_count();
    size_t smem = (size_t) bins * sizeof(uint32_t);
    if (smem <= 200 * 1024) {
        if (smem > 48 * 1024)
            B200_CUDA_CHECK(cudaFuncSetAttribute(histogram_smem_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int) smem));
        uint32_t per_sm = smem <= 32 * 1024 ? 2 : 1;
        uint32_t grid = (uint32_t) std::max<uint64_t>(
            1, std::min<uint64_t>((uint64_t) sms * per_sm, ceil_div(size, 4096)));
        histogram_smem_kernel<<<grid, 1024, smem, stream>>>(keys, size, bins, hist);
    } else {
        uint32_t grid = (uint32_t) std::max<uint64_t>(
            1, std::min<uint64_t>((uint64_t) sms * 8, ceil_div(size, 256)));
        histogram_global_kernel<<<grid, 256, 0, stream>>>(keys, size, bins, hist);
    }
    B200_LAUNCH_CHECK();
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

int b200_block_mkperm(void *stream_, const uint32_t *values, uint32_t size,
                      uint32_t block_size, uint32_t bucket_count, uint32_t *perm,
                      uint32_t *offsets, uint32_t *unique) {
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

    cudaStream_t stream = resolve_stream(stream_);
    const int sms = sm_count();
    const uint64_t ngroups = ceil_div(size, block_size);

    // ---- digit plan
    uint32_t total_bits = 0;
    while (total_bits < 32 && (1ull << total_bits) < bucket_count)
        total_bits++;
    const bool tiled = ngroups == 1; // one sorting group: tile kernels
    const uint32_t digit_bits = tiled ? 10 : 11;
    uint32_t npasses = 1, bits_per = 0;
    if (bucket_count > (tiled ? MT_MAX_BINS : MKPERM_MAX_BINS)) {
        npasses = (total_bits + digit_bits - 1) / digit_bits;
        bits_per = (total_bits + npasses - 1) / npasses;
    }
    const uint32_t ntiles = (uint32_t) ceil_div(size, MT_TILE);

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
    if (tiled) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_place_kernel<false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) tile_place_smem(MT_MAX_BINS)));
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_tile_place_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int) tile_place_smem(MT_MAX_BINS)));
    } else if (smem > 48 * 1024) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_hist_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        B200_CUDA_CHECK(cudaFuncSetAttribute(mkperm_place_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }

    // groups per launch so that the count table stays below 1 GiB
    uint64_t counts_per_group = (uint64_t) max_bins * (tiled ? ntiles : rows_per_group);
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

        if (tiled) {
            const uint64_t ncounts = (uint64_t) g.bins * ntiles;
            mkperm_tile_hist_kernel<<<(uint32_t) ceil_div(ntiles, MT_GROUP), MT_THREADS,
                                      (size_t) MT_GROUP * g.bins * 4, stream>>>(
                keys_in, size, ntiles, g.shift, g.mask, g.bins, counts);
            count_launch();
            rc = b200_block_prefix_reduce(stream, B200_VT_UINT32, B200_OP_ADD, ncounts, ncounts, 1, 0,
                                          counts, counts);
            if (rc) {
                cleanup();
                return rc;
            }
            // single pass: bucket starts are a strided view of the scanned table
            if (last && npasses == 1 && offsets) {
                uint32_t *records = (uint32_t *) temp_alloc(((size_t) bucket_count * 4 + 1) * 4, stream);
                if (!records) {
                    cleanup();
                    return fail(B200_ERR_CUDA, "jit_block_mkperm(): out of memory");
                }
                mkperm_offsets_kernel<<<1, 1024, 0, stream>>>(counts, ntiles, bucket_count, size, records);
                count_launch();
                cudaError_t err = cudaMemcpyAsync(offsets, records, ((size_t) bucket_count * 4 + 1) * 4,
                                                  cudaMemcpyDeviceToHost, stream);
                temp_free(records, stream);
                if (err != cudaSuccess) {
                    cleanup();
                    return cuda_fail(err, "cudaMemcpyAsync(offsets)");
                }
            }
            if (idx_in)
                mkperm_tile_place_kernel<true><<<ntiles, MT_THREADS, tile_place_smem(g.bins), stream>>>(
                    keys_in, idx_in, counts, size, ntiles, g.shift, g.mask, g.bins, idx_out, keys_out);
            else
                mkperm_tile_place_kernel<false><<<ntiles, MT_THREADS, tile_place_smem(g.bins), stream>>>(
                    keys_in, idx_in, counts, size, ntiles, g.shift, g.mask, g.bins, idx_out, keys_out);
            count_launch();
            continue;
        }

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
                cudaError_t err = cudaMemcpyAsync(offsets, records, ((size_t) bucket_count * 4 + 1) * 4,
                                                  cudaMemcpyDeviceToHost, stream);
                temp_free(records, stream);
                if (err != cudaSuccess) {
                    cleanup();
                    return cuda_fail(err, "cudaMemcpyAsync(offsets)");
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
            rc = cuda_fail(cudaMemcpyAsync(offsets, records, ((size_t) bucket_count * 4 + 1) * 4,
                                           cudaMemcpyDeviceToHost, stream),
                           "cudaMemcpyAsync(offsets)");
        }
        temp_free(hist, stream);
        if (rc) {
            cleanup();
            return rc;
        }
    }

    cleanup();

    if (offsets && ngroups == 1) {
        // the reference waits on an event here (src/cuda_ts.cpp:964-967)
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (unique)
            *unique = offsets[4 * (size_t) bucket_count];
    }
    return B200_OK;
}

} // extern "C"
