// sharded.cu -- large reductions, whole-array prefix reductions and the mkperm
// histogram sharded over the GPUs of one box: ONE C-ABI call per rank and
// primitive, totals exchanged through PEER-MAPPED MAILBOXES over NVLink.
//
// The reference has no multi-GPU code on this path (one CUDADevice + stream per
// GPU, src/cuda_core.cpp:352-516; no NCCL / MPI anywhere) -- this is the new
// functionality BASELINE.json's north_star asks for.  Data layout: rank r owns a
// contiguous shard of the global array; one process per GPU.
//
// Exchange: every rank allocates a small mailbox in its own device memory and
// exports it as a CUDA IPC handle; the host program passes the W handles around
// once (torch.distributed / any channel) and every rank maps all of them.  A
// collective then needs NO communication library call and no extra launch for the
// transfer: the last kernel of the local pass stores {value, epoch} straight into
// the W peers' mailboxes (plain st.global on peer addresses -> NVLink), and the
// consumer spins on ITS OWN mailbox (local L2) until W entries carry the epoch of
// this call.  Mailboxes are double buffered by epoch parity: a rank can be at most
// one collective ahead of its slowest peer, because every collective waits for all
// ranks (see b200_sharded_* in include/drjit_b200.h for the stream contract).
//
//   reduce     local reduce (reduce.cu) -> exchange kernel: post, wait, combine the W
//              partials in rank order (bit-exact integers, the same fixed floating
//              point order on every rank).
//   scan       tile sums of the shard (block_reduce with the 32 KiB scan tile) ->
//              seed kernel (one CTA): total of the tile sums, post, wait, carry of this
//              rank = totals of the ranks in front of it, seeds = carry (+) exclusive
//              scan of the tile sums -> seeded streaming scan (scan_fast.cu), which has
//              no look-back chain.  3 launches, no allocation, 12 B/element of HBM
//              traffic per GPU (4 for the sums, 8 for the scan).
//   histogram  local bucket counts -> exchange kernel: copy them into every peer's
//              mailbox, wait, sum the W rows (optionally also the rows of the lower
//              ranks = this rank's output offset per bucket).
#include "common.cuh"
#include "scan.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace b200 {

static constexpr uint32_t SH_MAX_WORLD = 16;
static constexpr uint32_t SH_MAX_BUCKETS = 65536;  // histogram rows that fit the mailbox
static constexpr uint64_t SH_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;
static constexpr uint32_t SH_CYC_ROUNDS = 4096;    // blocks per rank of the block-cyclic scan

/// Mailbox of one rank (device memory of that rank, mapped by every peer).
/// [parity][sender]: a 16-byte {value, epoch} slot and a histogram row.
struct Mailbox {
    struct Slot {
        uint64_t value;
        uint64_t epoch;
    };
    Slot scalar[2][SH_MAX_WORLD];
    Slot hist_flag[2][SH_MAX_WORLD];
    uint32_t hist[2][SH_MAX_WORLD][SH_MAX_BUCKETS];
    // block-cyclic scan: {total, epoch} of block (round, sender); entries are told apart
    // by the epoch of the call, so the table is never cleared
    Slot cyc[SH_CYC_ROUNDS][SH_MAX_WORLD];
};

} // namespace b200

struct B200Sharded {
    int rank = 0, world = 1, device = 0;
    b200::Mailbox *mine = nullptr;
    b200::Mailbox *peers[b200::SH_MAX_WORLD] = {};
    bool mapped[b200::SH_MAX_WORLD] = {};
    uint64_t epoch = 0;
    // scratch, grown on demand (never shrunk): tile sums and tile seeds of the scan,
    // partials, the error word of the bounded spin
    void *tsums = nullptr, *seeds = nullptr;
    size_t tile_capacity = 0; // bytes of each of the two
    uint64_t *scalars = nullptr; // [0] partial / total
    uint64_t *error = nullptr;   // pinned, mapped: set by a kernel whose wait for a peer timed out
    uint32_t *hist_local = nullptr; // SH_MAX_BUCKETS
    uint64_t *blockoff = nullptr;   // block-cyclic scan: {offset, epoch} per local block
};

namespace b200 {

struct PeerTable {
    Mailbox *box[SH_MAX_WORLD];
};

B200_DEVICE uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

B200_DEVICE void st_release_sys_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

B200_DEVICE uint64_t ld_acquire_sys_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/// Post {value, epoch} into slot[sender] of a (peer) mailbox: value first, then the
/// epoch with release semantics at system scope (the peer polls the epoch).
B200_DEVICE void post_slot(Mailbox::Slot *slot, uint64_t value, uint64_t epoch) {
    *(volatile uint64_t *) &slot->value = value;
    st_release_sys_u64(&slot->epoch, epoch);
}

/// Wait until the slot carries `epoch` (bounded: a peer that never arrives must not
/// hang the GPU); returns false on time-out.
B200_DEVICE bool wait_slot(const Mailbox::Slot *slot, uint64_t epoch, uint64_t &value) {
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys_u64(&slot->epoch) != epoch) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > SH_TIMEOUT_NS)
            return false;
        __nanosleep(64);
    }
    value = *(const volatile uint64_t *) &slot->value;
    return true;
}

template <typename V> B200_DEVICE uint64_t to_bits(V v) {
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(V));
    return b;
}

template <typename V> B200_DEVICE V from_bits(uint64_t b) {
    V v;
    memcpy(&v, &b, sizeof(V));
    return v;
}

/// reduce: post this rank's partial, wait for all, combine in rank order.
template <typename T, int Op>
__global__ void __launch_bounds__(32)
shard_reduce_exchange_kernel(const PeerTable peers, Mailbox *mine, uint32_t rank, uint32_t world,
                             uint64_t epoch, const T *partial, T *out, uint64_t *error) {
    using V = typename ValueOf<T>::type;
    using R = Red<V, Op>;
    const uint32_t lane = threadIdx.x, par = (uint32_t) (epoch & 1);
    const V mine_v = to_value<T>(*partial);
    if (lane < world)
        post_slot(&peers.box[lane]->scalar[par][rank], to_bits<V>(mine_v), epoch);
    uint64_t bits = 0;
    bool ok = true;
    if (lane < world)
        ok = wait_slot(&mine->scalar[par][lane], epoch, bits);
    if (!__all_sync(FULL_MASK, ok)) {
        if (lane == 0)
            *error = 1;
        return;
    }
    V acc = R::identity();
    for (uint32_t r = 0; r < world; ++r) // fixed order: rank 0 first
        acc = R::apply(acc, from_bits<V>(__shfl_sync(FULL_MASK, bits, r)));
    if (lane == 0)
        *out = from_value<T>(acc);
}

/// scan: one CTA.  total of the shard's tile sums -> post -> wait -> carry of this rank
/// -> seeds[t] = carry (+) tile sums in front of tile t (behind it when reverse).
template <typename V, int Op>
__global__ void __launch_bounds__(1024)
shard_scan_seed_kernel(const PeerTable peers, Mailbox *mine, uint32_t rank, uint32_t world,
                       uint64_t epoch, const V *__restrict__ tsums, V *__restrict__ seeds,
                       uint32_t ntiles, int reverse, uint64_t *error) {
    using R = Red<V, Op>;
    __shared__ V s_warp[32];
    __shared__ V s_carry;
    __shared__ int s_ok;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, par = (uint32_t) (epoch & 1);
    // thread t owns the contiguous logical range [t * per, (t + 1) * per) of the tile
    // sums; logical index l <-> physical tile (reverse ? ntiles - 1 - l : l)
    const uint32_t per = (ntiles + 1023u) / 1024u;
    const uint32_t l0 = min(tid * per, ntiles), l1 = min(l0 + per, ntiles);
    auto phys = [&](uint32_t l) { return reverse ? ntiles - 1 - l : l; };

    // (eight independent loads in flight: a dependent load per step would cost one L2 round
    // trip per tile sum -- 64 of them per thread for a 2 GiB shard)
    V local = R::identity();
    {
        uint32_t l = l0;
        for (; l + 8 <= l1; l += 8) {
            V v[8];
            #pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldg(tsums + phys(l + u));
            #pragma unroll
            for (int u = 0; u < 8; ++u)
                local = R::apply(local, v[u]);
        }
        for (; l < l1; ++l)
            local = R::apply(local, __ldg(tsums + phys(l)));
    }
    // block-wide inclusive scan of the thread totals (fixed order)
    V incl = local;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const V up = shfl_up(incl, d);
        if (lane >= (uint32_t) d)
            incl = R::apply(up, incl);
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    V wprefix = R::identity(), total = R::identity();
    for (uint32_t w = 0; w < 32; ++w) {
        const V t = s_warp[w];
        if (w < warp)
            wprefix = R::apply(wprefix, t);
        total = R::apply(total, t);
    }
    V excl = shfl_up(incl, 1);
    if (lane == 0)
        excl = R::identity();
    excl = R::apply(wprefix, excl); // tiles of this shard in front of the thread's range

    // ---- exchange of the W shard totals (warp 0)
    if (warp == 0) {
        if (lane < world)
            post_slot(&peers.box[lane]->scalar[par][rank], to_bits<V>(total), epoch);
        uint64_t bits = 0;
        bool ok = true;
        if (lane < world)
            ok = wait_slot(&mine->scalar[par][lane], epoch, bits);
        ok = __all_sync(FULL_MASK, ok);
        // carry = totals of the ranks in front of this one in scan direction
        V carry = R::identity();
        if (reverse) {
            for (uint32_t r = world; r-- > rank + 1;)
                carry = R::apply(carry, from_bits<V>(__shfl_sync(FULL_MASK, bits, r)));
        } else {
            for (uint32_t r = 0; r < rank; ++r)
                carry = R::apply(carry, from_bits<V>(__shfl_sync(FULL_MASK, bits, r)));
        }
        if (lane == 0) {
            s_carry = carry;
            s_ok = ok;
            if (!ok)
                *error = 1;
        }
    }
    __syncthreads();
    if (!s_ok)
        return;
    V run = R::apply(s_carry, excl);
    {
        uint32_t l = l0;
        for (; l + 8 <= l1; l += 8) {
            V v[8];
            #pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldg(tsums + phys(l + u));
            #pragma unroll
            for (int u = 0; u < 8; ++u) {
                seeds[phys(l + u)] = run;
                run = R::apply(run, v[u]);
            }
        }
        for (; l < l1; ++l) {
            const uint32_t p = phys(l);
            seeds[p] = run;
            run = R::apply(run, __ldg(tsums + p));
        }
    }
}

/// histogram: copy the local row into every peer's mailbox, wait, sum the rows.
__global__ void __launch_bounds__(1024)
shard_hist_exchange_kernel(const PeerTable peers, Mailbox *mine, uint32_t rank, uint32_t world,
                           uint64_t epoch, const uint32_t *__restrict__ local, uint32_t buckets,
                           uint32_t *__restrict__ global, uint32_t *__restrict__ before, uint64_t *error) {
    __shared__ int s_ok;
    const uint32_t tid = threadIdx.x, par = (uint32_t) (epoch & 1);
    for (uint32_t p = 0; p < world; ++p) {
        uint32_t *row = peers.box[p]->hist[par][rank];
        for (uint32_t b = tid; b < buckets; b += 1024)
            row[b] = local[b];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < 32) {
        if (tid < world)
            post_slot(&peers.box[tid]->hist_flag[par][rank], buckets, epoch);
        uint64_t seen = 0;
        bool ok = true;
        if (tid < world)
            ok = wait_slot(&mine->hist_flag[par][tid], epoch, seen) && seen == buckets;
        ok = __all_sync(FULL_MASK, ok);
        if (tid == 0) {
            s_ok = ok;
            if (!ok)
                *error = 1;
        }
    }
    __syncthreads();
    if (!s_ok)
        return;
    __threadfence_system();
    for (uint32_t b = tid; b < buckets; b += 1024) {
        uint32_t sum = 0, low = 0;
        for (uint32_t r = 0; r < world; ++r) {
            const uint32_t c = *(const volatile uint32_t *) &mine->hist[par][r][b];
            sum += c;
            low += r < rank ? c : 0u;
        }
        global[b] = sum;
        if (before)
            before[b] = low;
    }
}

typedef int (*ShardReduceFn)(B200Sharded *, cudaStream_t, const PeerTable &, const void *, void *);
typedef int (*ShardSeedFn)(B200Sharded *, cudaStream_t, const PeerTable &, uint32_t, int);

template <typename T, int Op>
static int launch_reduce_exchange(B200Sharded *c, cudaStream_t s, const PeerTable &pt, const void *partial,
                                  void *out) {
    shard_reduce_exchange_kernel<T, Op><<<1, 32, 0, s>>>(pt, c->mine, (uint32_t) c->rank, (uint32_t) c->world,
                                                         c->epoch, (const T *) partial, (T *) out,
                                                         c->error);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <typename V, int Op>
static int launch_seed(B200Sharded *c, cudaStream_t s, const PeerTable &pt, uint32_t ntiles, int reverse) {
    shard_scan_seed_kernel<V, Op><<<1, 1024, 0, s>>>(pt, c->mine, (uint32_t) c->rank, (uint32_t) c->world,
                                                     c->epoch, (const V *) c->tsums, (V *) c->seeds, ntiles,
                                                     reverse, c->error);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <typename T, bool Bits> static ShardReduceFn pick_reduce_exchange(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_reduce_exchange<T, B200_OP_ADD>;
        case B200_OP_MUL: return launch_reduce_exchange<T, B200_OP_MUL>;
        case B200_OP_MIN: return launch_reduce_exchange<T, B200_OP_MIN>;
        case B200_OP_MAX: return launch_reduce_exchange<T, B200_OP_MAX>;
        case B200_OP_AND: if constexpr (Bits) return launch_reduce_exchange<T, B200_OP_AND>; else return nullptr;
        case B200_OP_OR:  if constexpr (Bits) return launch_reduce_exchange<T, B200_OP_OR>; else return nullptr;
        default: return nullptr;
    }
}

template <typename V, bool Bits> static ShardSeedFn pick_seed(int op) {
    switch (op) {
        case B200_OP_ADD: return launch_seed<V, B200_OP_ADD>;
        case B200_OP_MUL: return launch_seed<V, B200_OP_MUL>;
        case B200_OP_MIN: return launch_seed<V, B200_OP_MIN>;
        case B200_OP_MAX: return launch_seed<V, B200_OP_MAX>;
        case B200_OP_AND: if constexpr (Bits) return launch_seed<V, B200_OP_AND>; else return nullptr;
        case B200_OP_OR:  if constexpr (Bits) return launch_seed<V, B200_OP_OR>; else return nullptr;
        default: return nullptr;
    }
}

static ShardReduceFn reduce_exchange_for(int vt, int op) {
    switch (vt) {
        case B200_VT_INT32:   return pick_reduce_exchange<int32_t, true>(op);
        case B200_VT_UINT32:  return pick_reduce_exchange<uint32_t, true>(op);
        case B200_VT_INT64:   return pick_reduce_exchange<int64_t, true>(op);
        case B200_VT_UINT64:  return pick_reduce_exchange<uint64_t, true>(op);
        case B200_VT_FLOAT16: return pick_reduce_exchange<__half, false>(op);
        case B200_VT_FLOAT32: return pick_reduce_exchange<float, false>(op);
        case B200_VT_FLOAT64: return pick_reduce_exchange<double, false>(op);
        default: return nullptr;
    }
}

static ShardSeedFn seed_for(int vt, int op) {
    switch (vt) {
        case B200_VT_INT32:   return pick_seed<int32_t, true>(op);
        case B200_VT_UINT32:  return pick_seed<uint32_t, true>(op);
        case B200_VT_INT64:   return pick_seed<int64_t, true>(op);
        case B200_VT_UINT64:  return pick_seed<uint64_t, true>(op);
        case B200_VT_FLOAT32: return pick_seed<float, false>(op);
        case B200_VT_FLOAT64: return pick_seed<double, false>(op);
        default: return nullptr;
    }
}

static int check_ctx(B200Sharded *c, const char *what) {
    if (!c || !c->mine)
        return fail(B200_ERR_INVALID, "%s: invalid context", what);
    for (int r = 0; r < c->world; ++r)
        if (!c->peers[r])
            return fail(B200_ERR_INVALID, "%s: rank %d is not connected (b200_sharded_connect)", what, r);
    // a peer that never arrived in an earlier collective leaves the error word set
    if (*(volatile uint64_t *) c->error)
        return fail(B200_ERR_CUDA, "%s: an earlier sharded call timed out waiting for a peer", what);
    return B200_OK;
}

static PeerTable peer_table(const B200Sharded *c) {
    PeerTable pt{};
    for (int r = 0; r < c->world; ++r)
        pt.box[r] = c->peers[r];
    return pt;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_sharded_create(int rank, int world, B200Sharded **out) {
    int rc = ensure_init();
    if (rc)
        return rc;
    if (!out || world < 1 || world > (int) SH_MAX_WORLD || rank < 0 || rank >= world)
        return fail(B200_ERR_INVALID, "b200_sharded_create(): invalid rank %d / world size %d (at most %u ranks)",
                    rank, world, SH_MAX_WORLD);
    B200Sharded *c = new B200Sharded();
    c->rank = rank;
    c->world = world;
    cudaGetDevice(&c->device);
    // plain cudaMalloc: pool memory cannot be exported through CUDA IPC
    cudaError_t err = cudaMalloc((void **) &c->mine, sizeof(Mailbox));
    if (err == cudaSuccess)
        err = cudaMemset(c->mine, 0, sizeof(Mailbox));
    if (err == cudaSuccess)
        err = cudaMalloc((void **) &c->scalars, 64);
    if (err == cudaSuccess)
        err = cudaMemset(c->scalars, 0, 64);
    if (err == cudaSuccess)
        err = cudaMalloc((void **) &c->hist_local, SH_MAX_BUCKETS * sizeof(uint32_t));
    if (err == cudaSuccess)
        err = cudaMalloc((void **) &c->blockoff, SH_CYC_ROUNDS * 16);
    if (err == cudaSuccess)
        err = cudaMemset(c->blockoff, 0, SH_CYC_ROUNDS * 16);
    if (err == cudaSuccess)
        err = cudaHostAlloc((void **) &c->error, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) {
        memset(c->error, 0, 64);
        err = cudaDeviceSynchronize(); // the mailbox is zero before its handle leaves this process
    }
    if (err != cudaSuccess) {
        b200_sharded_destroy(c);
        return cuda_fail(err, "b200_sharded_create()");
    }
    c->peers[rank] = c->mine;
    *out = c;
    return B200_OK;
}

int b200_sharded_handle_bytes(void) { return (int) sizeof(cudaIpcMemHandle_t); }

int b200_sharded_export(B200Sharded *c, void *handle) {
    if (!c || !c->mine || !handle)
        return fail(B200_ERR_INVALID, "b200_sharded_export(): invalid argument");
    cudaIpcMemHandle_t h;
    B200_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->mine));
    memcpy(handle, &h, sizeof(h));
    return B200_OK;
}

int b200_sharded_connect(B200Sharded *c, const void *handles) {
    if (!c || !c->mine || !handles)
        return fail(B200_ERR_INVALID, "b200_sharded_connect(): invalid argument");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->peers[r])
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *) handles + (size_t) r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess)
            return cuda_fail(err, "b200_sharded_connect(): cudaIpcOpenMemHandle");
        c->peers[r] = (Mailbox *) p;
        c->mapped[r] = true;
    }
    return B200_OK;
}

int b200_sharded_destroy(B200Sharded *c) {
    if (!c)
        return B200_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->mapped[r] && c->peers[r])
            cudaIpcCloseMemHandle(c->peers[r]);
    cudaFree(c->mine);
    cudaFree(c->scalars);
    cudaFree(c->hist_local);
    cudaFree(c->blockoff);
    if (c->error)
        cudaFreeHost(c->error);
    cudaFree(c->tsums);
    cudaFree(c->seeds);
    cudaGetLastError();
    delete c;
    return B200_OK;
}

int b200_sharded_reduce(B200Sharded *c, void *stream_, int vt, int op, const void *in,
                        uint64_t local_size, void *out) {
    int rc = check_ctx(c, "b200_sharded_reduce()");
    if (rc)
        return rc;
    ShardReduceFn fn = reduce_exchange_for(vt, op);
    if (!fn)
        return fail(B200_ERR_UNSUPPORTED, "b200_sharded_reduce(): no existing kernel for type=%s, op=%s!",
                    type_name(vt), op_name(op));
    cudaStream_t stream = resolve_stream(stream_);
    void *partial = c->scalars;
    if (local_size > 0) {
        if ((rc = b200_reduce(stream, vt, op, in, local_size, partial)))
            return rc;
    } else {
        uint64_t ident = b200_reduce_identity(vt, op);
        B200_CUDA_CHECK(cudaMemcpyAsync(partial, &ident, 8, cudaMemcpyHostToDevice, stream));
    }
    c->epoch++;
    return fn(c, stream, peer_table(c), partial, out);
}

int b200_sharded_reduce_dot(B200Sharded *c, void *stream_, int vt, const void *a, const void *b,
                            uint64_t local_size, void *out) {
    int rc = check_ctx(c, "b200_sharded_reduce_dot()");
    if (rc)
        return rc;
    ShardReduceFn fn = vt == B200_VT_FLOAT16 || vt == B200_VT_FLOAT32 || vt == B200_VT_FLOAT64
                           ? reduce_exchange_for(vt, B200_OP_ADD) : nullptr;
    if (!fn)
        return fail(B200_ERR_UNSUPPORTED, "b200_sharded_reduce_dot(): no existing kernel for type=%s!", type_name(vt));
    cudaStream_t stream = resolve_stream(stream_);
    void *partial = c->scalars;
    if (local_size > 0) {
        if ((rc = b200_reduce_dot(stream, vt, a, b, local_size, partial)))
            return rc;
    } else {
        B200_CUDA_CHECK(cudaMemsetAsync(partial, 0, 8, stream));
    }
    c->epoch++;
    return fn(c, stream, peer_table(c), partial, out);
}

int b200_sharded_prefix_reduce(B200Sharded *c, void *stream_, int vt, int op, uint64_t local_size,
                               int exclusive, int reverse, const void *in, void *out) {
    int rc = check_ctx(c, "b200_sharded_prefix_reduce()");
    if (rc)
        return rc;
    ShardSeedFn fn = seed_for(vt, op);
    if (!fn)
        return fail(B200_ERR_UNSUPPORTED,
                    "b200_sharded_prefix_reduce(): no existing kernel for type=%s, op=%s!", type_name(vt),
                    op_name(op));
    if (((uintptr_t) in | (uintptr_t) out) & 15)
        return fail(B200_ERR_INVALID, "b200_sharded_prefix_reduce(): the shard must be 16-byte aligned!");
    const uint32_t tsize = type_size(vt), tile = b200_scan_tile_elems(vt);
    if (local_size >= 0xffffffffull - 2 * tile)
        return fail(B200_ERR_INVALID, "b200_sharded_prefix_reduce(): shard too large (< 2^32 - 2^14 elements)!");
    cudaStream_t stream = resolve_stream(stream_);
    const uint32_t ntiles = (uint32_t) ceil_div(local_size, tile);
    const size_t need = std::max<size_t>((size_t) ntiles * tsize, 64);
    if (need > c->tile_capacity) {
        // (grows rarely; a synchronising cudaFree is fine here)
        cudaFree(c->tsums);
        cudaFree(c->seeds);
        c->tsums = c->seeds = nullptr;
        c->tile_capacity = 0;
        size_t cap = 1;
        while (cap < need)
            cap <<= 1;
        B200_CUDA_CHECK(cudaMalloc(&c->tsums, cap));
        B200_CUDA_CHECK(cudaMalloc(&c->seeds, cap));
        c->tile_capacity = cap;
    }
    // pass 1 (4 B / element): one sum per scan tile; the last tile may be short
    if (local_size > 0 && (rc = b200_block_reduce(stream, vt, op, local_size, std::min<uint64_t>(tile, local_size),
                                                   in, c->tsums)))
        return rc;
    // exchange + seeds (one CTA)
    c->epoch++;
    if ((rc = fn(c, stream, peer_table(c), ntiles, reverse != 0)))
        return rc;
    // pass 2 (8 B / element): every tile scans on its own, seeded with its prefix
    if (local_size > 0)
        rc = b200_prefix_reduce_seeded(stream, vt, op, local_size, exclusive, reverse, in, out, c->seeds);
    return rc;
}

int b200_sharded_prefix_reduce_cyclic(B200Sharded *c, void *stream_, int vt, int op, uint64_t local_size,
                                      uint64_t block_size, int exclusive, const void *in, void *out) {
    int rc = check_ctx(c, "b200_sharded_prefix_reduce_cyclic()");
    if (rc)
        return rc;
    const uint32_t tile = b200_scan_tile_elems(vt);
    if (!seed_for(vt, op) || tile == 0)
        return fail(B200_ERR_UNSUPPORTED,
                    "b200_sharded_prefix_reduce_cyclic(): no existing kernel for type=%s, op=%s!", type_name(vt),
                    op_name(op));
    if (((uintptr_t) in | (uintptr_t) out) & 15)
        return fail(B200_ERR_INVALID, "b200_sharded_prefix_reduce_cyclic(): the shard must be 16-byte aligned!");
    if (block_size < tile || (block_size & (block_size - 1)) != 0 || local_size % block_size != 0 ||
        local_size / block_size > SH_CYC_ROUNDS || local_size >= 0xffffffffull - 2 * tile)
        return fail(B200_ERR_INVALID,
                    "b200_sharded_prefix_reduce_cyclic(): blocks must be a power of two of at least %u elements, "
                    "shards whole blocks (at most %u, fewer than 2^32 - 2^14 elements)!", tile, SH_CYC_ROUNDS);
    if (local_size == 0)
        return B200_OK;
    cudaStream_t stream = resolve_stream(stream_);
    c->epoch++;
    CyclicScan cyc{};
    for (int r = 0; r < c->world; ++r)
        cyc.table[r] = (uint64_t *) &c->peers[r]->cyc[0][0];
    cyc.blockoff = c->blockoff;
    cyc.error = c->error;
    cyc.epoch = c->epoch;
    cyc.rank = (uint32_t) c->rank;
    cyc.world = (uint32_t) c->world;
    cyc.log2_block_tiles = log2i(block_size / tile);
    cyc.table_stride = SH_MAX_WORLD;
    ScanCall call{ stream, in, out, local_size, local_size, exclusive != 0, false, nullptr, nullptr, true };
    call.cyclic = &cyc;
    bool handled = false;
    // (vt / op canonicalised like b200_block_prefix_reduce: signed add / and / or on the unsigned kernels)
    rc = scan_fast_dispatch(vt, op, call, &handled);
    if (rc)
        return rc;
    if (!handled)
        return fail(B200_ERR_UNSUPPORTED, "b200_sharded_prefix_reduce_cyclic(): unsupported configuration");
    return B200_OK;
}

int b200_sharded_histogram(B200Sharded *c, void *stream_, const uint32_t *values, uint64_t local_size,
                           uint32_t bucket_count, uint32_t *hist, uint32_t *before) {
    int rc = check_ctx(c, "b200_sharded_histogram()");
    if (rc)
        return rc;
    if (bucket_count == 0 || bucket_count > SH_MAX_BUCKETS)
        return fail(B200_ERR_INVALID, "b200_sharded_histogram(): bucket_count must be in 1..%u!", SH_MAX_BUCKETS);
    cudaStream_t stream = resolve_stream(stream_);
    if ((rc = b200_mkperm_histogram(stream, values, local_size, bucket_count, c->hist_local)))
        return rc;
    c->epoch++;
    shard_hist_exchange_kernel<<<1, 1024, 0, stream>>>(peer_table(c), c->mine, (uint32_t) c->rank,
                                                       (uint32_t) c->world, c->epoch, c->hist_local,
                                                       bucket_count, hist, before, c->error);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

} // extern "C"
