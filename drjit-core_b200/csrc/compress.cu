// compress.cu -- stream compaction of a byte mask into ascending indices.
//
// Replaces CUDAThreadState::compress (src/cuda_ts.cpp:683-763) and
// resources/compress.cuh (compress_small / compress_large_init /
// compress_large).  One kernel for every size:
//   - 256 threads x J 16-byte vectors = 16 KiB of mask per tile (the reference:
//     2 KiB), tiles handed out by an atomic ticket and chained with decoupled
//     look-back on 64-bit {status, count} descriptors;
//   - the mask is never written: the tail and a misaligned head are handled by
//     guarded loads instead of the reference's memset of the trailer
//     (src/cuda_ts.cpp:708-710, :746-748);
//   - indices are compacted per warp row in shared memory and written with
//     contiguous (coalesced) stores instead of one predicated scattered store
//     per mask byte (compress.cuh:146-149).
#include "common.cuh"


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
