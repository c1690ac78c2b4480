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
    jit_shutdown(0);
    printf("jit_h_client: %d checks, %d failed\n", tests, failures);
    return failures ? 1 : 0;
}
