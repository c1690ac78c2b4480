// pipeline.cuh -- Blackwell/Hopper asynchronous-copy building blocks used by the
// streaming kernels: mbarrier, 1-D bulk copies (the TMA engine without a tensor
// map: cp.async.bulk global <-> shared) and the proxy fence between generic
// shared-memory writes and the async proxy.
//
// The reference has nothing comparable (its kernels are compute_75 PTX: one
// ld.global per thread, resources/block_prefix_reduce.cuh:112-131).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define B200_PIPE_DEVICE __device__ __forceinline__

namespace b200 {

B200_PIPE_DEVICE uint32_t smem_addr(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}

B200_PIPE_DEVICE void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}

/// Make mbarrier initialisation visible to the async proxy (before the first bulk copy)
B200_PIPE_DEVICE void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

/// One arrival that also announces 'bytes' of asynchronous-copy traffic
B200_PIPE_DEVICE void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}

B200_PIPE_DEVICE void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}

/// Block until the phase with the given parity has completed
B200_PIPE_DEVICE void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_addr(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" :: "r"(addr), "r"(parity) : "memory");
}

/// global -> shared bulk copy; completion is signalled on 'bar' (complete_tx).
/// Addresses 16-byte aligned, 'bytes' a multiple of 16.
B200_PIPE_DEVICE void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

/// shared -> global bulk copy, tracked by the thread's bulk async-group
B200_PIPE_DEVICE void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}

B200_PIPE_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

/// Wait until at most N of the thread's bulk groups still READ shared memory
template <int N> B200_PIPE_DEVICE void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}

/// Wait until at most N of the thread's bulk groups are incomplete
template <int N> B200_PIPE_DEVICE void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}

/// Order generic-proxy shared-memory writes before async-proxy reads (bulk stores)
B200_PIPE_DEVICE void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace b200
