// ubench_warp_ops.cu -- development aid: issue cost of the warp-level operations
// the mkperm / scatter kernels are built from, on the GPU at hand.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_warp_ops ubench_warp_ops.cu
// Every test runs ITER dependent-free operations per warp in 32 warps per SM on
// every SM and reports cycles per warp-instruction per SM sub-partition (SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define THREADS 1024

enum { MATCH32, MATCH16, MATCH4, MATCH1, BALLOT, SHFL, ATOMS_ADD32, ATOMS_ADD16, ATOMS_ADD4, ATOMS_ADD1,
       ATOMS_OR32, ATOMS_OR4, REDUX, LDS_STS, POPC, ATOMS_NORET32, ATOMS_NORET_STRIDED, LDS_ONLY, STS_ONLY, LDS128_ONLY, NTESTS };
static const char *names[NTESTS] = { "match.any 32 distinct", "match.any 16 distinct", "match.any 4 distinct",
    "match.any 1 distinct", "vote.ballot", "shfl.idx", "atoms.add(ret) 32 addr", "atoms.add(ret) 16 addr",
    "atoms.add(ret) 4 addr", "atoms.add(ret) 1 addr", "atoms.or(noret) 32 addr", "atoms.or(noret) 4 addr",
    "redux.add", "lds+sts private", "popc", "atoms.add(noret) 32 addr", "atoms.add(noret) own column", "lds independent",
    "sts independent", "lds.128 independent" };

template <int TEST> __global__ void __launch_bounds__(THREADS) bench(uint32_t *out, long long *cycles) {
    __shared__ uint32_t s[32][64];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 64; i += 32)
        s[warp][i] = 0;
    __syncthreads();
    uint32_t acc = 0, v = lane * 2654435761u + warp;
    long long t0 = clock64();
    #pragma unroll 8
    for (int i = 0; i < ITER; ++i) {
        uint32_t x = v + i;
        if (TEST == MATCH32) acc += __match_any_sync(0xffffffffu, (x & ~31u) | lane);
        if (TEST == MATCH16) acc += __match_any_sync(0xffffffffu, (x & ~31u) | (lane & 15));
        if (TEST == MATCH4) acc += __match_any_sync(0xffffffffu, (x & ~31u) | (lane & 3));
        if (TEST == MATCH1) acc += __match_any_sync(0xffffffffu, i);
        if (TEST == BALLOT) acc += __ballot_sync(0xffffffffu, (x >> (i & 7)) & 1);
        if (TEST == SHFL) acc += __shfl_sync(0xffffffffu, x, (i + lane) & 31);
        if (TEST == ATOMS_ADD32) acc += atomicAdd(&s[warp][lane], 1u);
        if (TEST == ATOMS_ADD16) acc += atomicAdd(&s[warp][lane & 15], 1u);
        if (TEST == ATOMS_ADD4) acc += atomicAdd(&s[warp][lane & 3], 1u);
        if (TEST == ATOMS_ADD1) acc += atomicAdd(&s[warp][0], 1u);
        if (TEST == ATOMS_OR32) atomicOr(&s[warp][(lane + i) & 31], 1u << lane);
        if (TEST == ATOMS_OR4) atomicOr(&s[warp][(lane + i) & 3], 1u << lane);
        if (TEST == REDUX) acc += __reduce_add_sync(0xffffffffu, x);
        if (TEST == LDS_STS) { uint32_t o = s[warp][(lane + i) & 63]; s[warp][(lane + i) & 63] = o + x; acc += o; }
        if (TEST == POPC) acc += __popc(x);
        if (TEST == ATOMS_NORET32) atomicAdd(&s[warp][lane + (i & 32)], 1u);
        if (TEST == ATOMS_NORET_STRIDED) atomicAdd(&((uint32_t *) s)[((x >> 7) & 15) * 128 + (threadIdx.x & 127)], 1u << (16 * (warp >> 4)));
        if (TEST == LDS_ONLY) acc += s[warp][(lane + i) & 63];
        if (TEST == STS_ONLY) s[warp][(lane + i) & 63] = x;
        if (TEST == LDS128_ONLY) { uint4 q = ((const uint4 *) s)[(warp * 16 + ((lane + i) & 15))]; acc += q.x + q.w; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *cycles = t1 - t0;
    if (acc == 0x12345678u)
        out[0] = acc + s[warp][lane];
}

template <int TEST> static void run(uint32_t *out, long long *cyc, int sms) {
    bench<TEST><<<sms, THREADS>>>(out, cyc);
    cudaDeviceSynchronize();
    bench<TEST><<<sms, THREADS>>>(out, cyc);
    cudaError_t err = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // 32 warps per SM = 8 per SMSP, ITER operations each
    printf("%-28s %8.2f cycles per warp-instruction per SMSP   (%s)\n", names[TEST],
           (double) c / ((double) ITER * 8.0), cudaGetErrorString(err));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, 64);
    cudaMalloc(&cyc, 8);
    printf("SMs: %d\n", sms);
    run<MATCH32>(out, cyc, sms); run<MATCH16>(out, cyc, sms); run<MATCH4>(out, cyc, sms); run<MATCH1>(out, cyc, sms);
    run<BALLOT>(out, cyc, sms); run<SHFL>(out, cyc, sms);
    run<ATOMS_ADD32>(out, cyc, sms); run<ATOMS_ADD16>(out, cyc, sms); run<ATOMS_ADD4>(out, cyc, sms);
    run<ATOMS_ADD1>(out, cyc, sms); run<ATOMS_OR32>(out, cyc, sms); run<ATOMS_OR4>(out, cyc, sms);
    run<REDUX>(out, cyc, sms); run<LDS_STS>(out, cyc, sms); run<POPC>(out, cyc, sms);
    run<ATOMS_NORET32>(out, cyc, sms); run<ATOMS_NORET_STRIDED>(out, cyc, sms); run<LDS_ONLY>(out, cyc, sms);
    run<STS_ONLY>(out, cyc, sms); run<LDS128_ONLY>(out, cyc, sms);
    return 0;
}
