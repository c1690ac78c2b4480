// tests/cpp/jit_h_client.cpp -- tier 2: a C++ translation unit written
// against the jit.h entry points of the path, linked with libdrjit_core_b200.so.
//
// Built twice (oracle/Makefile):
//   -DUSE_REFERENCE_HEADER  against the reference's OWN <drjit-core/jit.h> (only
//                           where /root/reference exists; the binary travels to
//                           the GPU box) -- i.e. an unmodified caller of the
//                           reference links and runs on this library;
//   (default)               against include/drjit_b200_jit.h, the mirror header.
// The checks follow tests/reductions.cpp of the reference (sizes 23 i^3 + 1, fmix32
// inputs, serial host loops as the expected values).  Exit code 0 = all passed.
#if defined(USE_REFERENCE_HEADER)
#  include <drjit-core/jit.h>
#else
#  include "drjit_b200_jit.h"
#endif

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

static uint32_t mix(uint32_t h) { // 32-bit murmur finaliser of i + 1
    h += 1;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAILED " __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

template <typename T> struct Dev {
    T *ptr;
    size_t n;
    explicit Dev(size_t n) : ptr((T *) jit_malloc(JitBackend::CUDA, std::max<size_t>(n, 1) * sizeof(T), 0)), n(n) {}
    ~Dev() { jit_free(ptr); }
    void put(const std::vector<T> &h) { jit_memcpy(JitBackend::CUDA, ptr, h.data(), h.size() * sizeof(T)); }
    std::vector<T> get(size_t count) const {
        std::vector<T> h(count);
        jit_sync_thread();
        jit_memcpy(JitBackend::CUDA, h.data(), ptr, count * sizeof(T));
        return h;
    }
};

int main() {
    jit_init(1u << (uint32_t) JitBackend::CUDA);
    if (!jit_has_backend(JitBackend::CUDA)) {
        fprintf(stderr, "no CUDA backend\n");
        return 2;
    }
    int tests = 0;
    for (uint32_t i = 0; i < 24; i += 3) {
        const uint32_t size = 23 * i * i * i + 1;
        std::vector<uint32_t> h(size);
        for (uint32_t k = 0; k < size; ++k)
            h[k] = mix(k);
        Dev<uint32_t> d_in(size), d_out(size);
        d_in.put(h);

        for (uint32_t j = 0; j < 24; j += 5) {
            const uint32_t bs = std::min(size, 23 * j * j * j + 1);
            const uint32_t nb = (size + bs - 1) / bs;
            // block reduce (tests/reductions.cpp:109-151)
            jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, size, bs, d_in.ptr, d_out.ptr);
            std::vector<uint32_t> got = d_out.get(nb);
            bool ok = true;
            for (uint32_t b = 0; b < nb && ok; ++b) {
                uint32_t s = 0;
                for (uint32_t k = b * bs; k < std::min(size, (b + 1) * bs); ++k)
                    s += h[k];
                ok = got[b] == s;
            }
            CHECK(ok, "block_reduce size=%u bs=%u", size, bs);
            tests++;
            // prefix sums (tests/reductions.cpp:153-267): positional (size, block_size)
            for (int excl = 0; excl < 2; ++excl)
                for (int rev = 0; rev < 2; ++rev) {
                    jit_block_prefix_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, size, bs, excl, rev,
                                            d_in.ptr, d_out.ptr);
                    got = d_out.get(size);
                    ok = true;
                    for (uint32_t b = 0; b < nb && ok; ++b) {
                        const uint32_t lo = b * bs, hi = std::min(size, (b + 1) * bs);
                        uint32_t s = 0;
                        for (uint32_t q = 0; q < hi - lo && ok; ++q) {
                            const uint32_t k = rev ? hi - 1 - q : lo + q;
                            if (excl) { ok = got[k] == s; s += h[k]; } else { s += h[k]; ok = got[k] == s; }
                        }
                    }
                    CHECK(ok, "block_prefix_reduce size=%u bs=%u excl=%d rev=%d", size, bs, excl, rev);
                    tests++;
                }
        }
        // jit_reduce as declared in jit.h:2219 (in, size, out)
        jit_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, (const void *) d_in.ptr, size, (void *) d_out.ptr);
        uint32_t total = 0;
        for (uint32_t v : h)
            total += v;
        CHECK(d_out.get(1)[0] == total, "reduce size=%u", size);
        tests++;

        // compress (tests/reductions.cpp:269-313)
        std::vector<uint8_t> m(size);
        std::vector<uint32_t> expect;
        for (uint32_t k = 0; k < size; ++k) {
            m[k] = (h[k] >> 9) & 1;
            if (m[k])
                expect.push_back(k);
        }
        // the reference requires the mask buffer to be padded (jit.h:2382-2383): jit_malloc rounds up
        Dev<uint8_t> d_m(size);
        d_m.put(m);
        const uint32_t count = jit_compress(JitBackend::CUDA, d_m.ptr, size, d_out.ptr);
        CHECK(count == expect.size(), "compress count size=%u", size);
        if (count == expect.size())
            CHECK(d_out.get(count) == expect, "compress indices size=%u", size);
        tests++;

        // mkperm (tests/reductions.cpp:315-406): per-bucket index sets
        for (uint32_t buckets : { 1u, 23u, 1000u, 70001u }) {
            std::vector<uint32_t> keys(size);
            for (uint32_t k = 0; k < size; ++k)
                keys[k] = h[k] % buckets;
            d_in.put(keys);
            uint32_t *offsets = (uint32_t *) jit_malloc(JitBackend::CUDA, (4 * (size_t) buckets + 1) * 4, 1);
            const uint32_t unique = jit_block_mkperm(JitBackend::CUDA, d_in.ptr, size, size, buckets, d_out.ptr, offsets);
            std::vector<uint32_t> perm = d_out.get(size);
            std::vector<std::vector<uint32_t>> sets(buckets);
            for (uint32_t k = 0; k < size; ++k)
                sets[keys[k]].push_back(k);
            uint32_t nonempty = 0;
            for (auto &s : sets)
                nonempty += !s.empty();
            bool ok = unique == nonempty && offsets[4 * (size_t) buckets] == unique;
            uint64_t covered = 0;
            for (uint32_t r = 0; r < unique && ok; ++r) {
                const uint32_t id = offsets[4 * r], start = offsets[4 * r + 1], cnt = offsets[4 * r + 2];
                ok = id < buckets && cnt == sets[id].size() && (uint64_t) start + cnt <= size;
                if (ok) {
                    std::vector<uint32_t> mine(perm.begin() + start, perm.begin() + start + cnt);
                    std::sort(mine.begin(), mine.end());
                    ok = mine == sets[id];
                    covered += cnt;
                }
            }
            CHECK(ok && covered == size, "mkperm size=%u buckets=%u", size, buckets);
            jit_free(offsets);
            tests++;
            d_in.put(h);
        }
    }
    // typed fill (jit_memset_async, jit.h:2199) and identities (jit.h:2840)
    {
        Dev<uint64_t> d(1000);
        const uint64_t pat = 0x0123456789abcdefull;
        jit_memset_async(JitBackend::CUDA, d.ptr, 1000, 8, &pat);
        std::vector<uint64_t> got = d.get(1000);
        CHECK(std::all_of(got.begin(), got.end(), [&](uint64_t v) { return v == pat; }), "memset_async");
        CHECK(jit_reduce_identity(VarType::UInt32, ReduceOp::Min) == 0xffffffffull, "identity u32 min");
        CHECK(jit_reduce_identity(VarType::Float32, ReduceOp::Mul) == 0x3f800000ull, "identity f32 mul");
        tests += 3;
    }
    // error behaviour: std::runtime_error across the API (src/log.cpp:165-169)
    {
        Dev<uint32_t> d(16);
        bool thrown = false;
        try {
            jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, 16, 0, d.ptr, d.ptr);
        } catch (const std::runtime_error &) {
            thrown = true;
        }
        CHECK(thrown, "block_size == 0 must throw");
        thrown = false;
        try {
            jit_block_reduce(JitBackend::LLVM, VarType::UInt32, ReduceOp::Add, 16, 4, d.ptr, d.ptr);
        } catch (const std::runtime_error &) {
            thrown = true;
        }
        CHECK(thrown, "non-CUDA backend must throw (no CPU fallback)");
        tests += 2;
    }
    // flags of the boundary (SURVEY.md 8b "Ordering / sync"; jit.h:1734-1742, :1815-1824)
    {
        const uint32_t n = 100000;
        std::vector<uint32_t> h(n);
        std::vector<uint8_t> hm(n);
        uint32_t total = 0, expect_count = 0;
        for (uint32_t k = 0; k < n; ++k) {
            h[k] = mix(k);
            total += h[k];
            hm[k] = (mix(k) & 7) == 0;
            expect_count += hm[k];
        }
        Dev<uint32_t> d_in(n), d_out(n);
        Dev<uint8_t> d_mask(2 * n);
        d_in.put(h);
        d_mask.put(hm);
        const uint32_t flags0 = jit_flags();

        // KernelHistory: every primitive call shows up with its KernelType
        jit_kernel_history_clear();
        jit_set_flag(JitFlag::KernelHistory, 1);
        CHECK(jit_flag(JitFlag::KernelHistory) == 1, "jit_flag(KernelHistory)");
        jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, d_in.ptr, d_out.ptr);
        jit_block_prefix_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, 1, 0, d_in.ptr, d_out.ptr);
        uint32_t cnt = jit_compress(JitBackend::CUDA, d_mask.ptr, n, d_out.ptr);
        jit_block_mkperm(JitBackend::CUDA, d_in.ptr, n, n, 0xffffffffu, d_out.ptr, nullptr);
        jit_set_flag(JitFlag::KernelHistory, 0);
        KernelHistoryEntry *hist = jit_kernel_history();
        CHECK(hist != nullptr, "kernel history is empty");
        bool seen[4] = { false, false, false, false };
        size_t entries = 0;
        for (KernelHistoryEntry *e = hist; e && (uint32_t) e->backend; ++e, ++entries) {
            // (mkperm scans its count tables through block_prefix_reduce, which adds
            // entries of other sizes -- as in the reference)
            CHECK(e->backend == JitBackend::CUDA && e->execution_time >= 0.f, "history entry");
            if (e->type == KernelType::BlockReduce && e->size == n) seen[0] = true;
            if (e->type == KernelType::BlockPrefixReduce && e->size == n) seen[1] = true;
            if (e->type == KernelType::Compress && e->size == n) seen[2] = true;
            if (e->type == KernelType::MkPerm && e->size == n) seen[3] = true;
            free(e->ir);
        }
        free(hist);
        CHECK(entries >= 4 && seen[0] && seen[1] && seen[2] && seen[3], "history types (%zu entries)", entries);
        CHECK(cnt == expect_count, "compress under KernelHistory");
        CHECK(jit_kernel_history() == nullptr, "jit_kernel_history() must clear the history");
        tests += 5;

        // LaunchBlocking: the result is there when the call returns
        jit_set_flag(JitFlag::LaunchBlocking, 1);
        uint32_t *pinned = (uint32_t *) jit_malloc(JitBackend::CUDA, 64, 1);
        pinned[0] = 0;
        jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, d_in.ptr, pinned);
        CHECK(pinned[0] == total, "LaunchBlocking: result must be visible without a sync");
        jit_set_flag(JitFlag::LaunchBlocking, 0);
        jit_free(pinned);
        tests++;

        // ForbidSynchronization: the synchronising entry points raise (src/init.cpp:503-505),
        // the asynchronous ones keep working
        jit_set_flag(JitFlag::ForbidSynchronization, 1);
        bool thrown = false;
        try {
            jit_compress(JitBackend::CUDA, d_mask.ptr, n, d_out.ptr);
        } catch (const std::runtime_error &) {
            thrown = true;
        }
        CHECK(thrown, "jit_compress must throw under ForbidSynchronization");
        thrown = false;
        try {
            jit_sync_thread();
        } catch (const std::runtime_error &) {
            thrown = true;
        }
        CHECK(thrown, "jit_sync_thread must throw under ForbidSynchronization");
        thrown = false;
        try {
            jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, d_in.ptr, d_out.ptr);
        } catch (const std::runtime_error &) {
            thrown = true;
        }
        CHECK(!thrown, "asynchronous primitives must not throw under ForbidSynchronization");
        jit_set_flag(JitFlag::ForbidSynchronization, 0);
        CHECK(d_out.get(1)[0] == total, "reduce enqueued under ForbidSynchronization");
        jit_set_flags(flags0);
        CHECK(jit_flags() == flags0, "jit_set_flags / jit_flags");
        tests += 5;

        // jit_malloc_migrate (jit.h:516): device -> host copy, host -> device move
        uint32_t *host = (uint32_t *) jit_malloc_migrate(d_in.ptr, JitBackend::None, 0);
        jit_sync_thread();
        CHECK(host && host != d_in.ptr && std::memcmp(host, h.data(), n * sizeof(uint32_t)) == 0,
              "jit_malloc_migrate to the host");
        uint32_t *dev2 = (uint32_t *) jit_malloc_migrate(host, JitBackend::CUDA, 1);
        jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, dev2, d_out.ptr);
        CHECK(d_out.get(1)[0] == total, "jit_malloc_migrate back to the device");
        CHECK(jit_malloc_migrate(dev2, JitBackend::CUDA, 1) == dev2, "migrate to the same place is a no-op");
        jit_free(dev2);
        tests += 3;

        // jit_cuda_sync_stream (jit.h:243-255): the per-thread default stream (2) waits
        jit_block_reduce(JitBackend::CUDA, VarType::UInt32, ReduceOp::Add, n, n, d_in.ptr, d_out.ptr);
        jit_cuda_sync_stream(2);
        tests++;
    }
    jit_shutdown(0);
    printf("jit_h_client: %d checks, %d failed\n", tests, failures);
    return failures ? 1 : 0;
}
