// scan.cuh -- interface between the prefix-reduction dispatcher (scan.cu, which
// also holds the fully general segmented kernel) and the streaming fast paths
// (scan_fast.cu).
#pragma once

#include "common.cuh"

namespace b200 {

/// Block-cyclic whole-array scan over the GPUs of one box (sharded.cu): the global array
/// is cut into blocks of 2^log2_block_tiles scan tiles, global block b lives on rank
/// b % world as that rank's local block b / world.  Every rank runs ONE chained
/// streaming scan over its local blocks; a block's tiles chain among themselves
/// (block-relative prefixes), its last tile posts the block total {value, epoch} into the
/// `table` of every rank (peer-mapped memory, one 16-byte store over NVLink), and its
/// first tile turns the totals of all blocks in front of it (global order) into the
/// block's offset, which it leaves in `blockoff` for the other tiles of the block.
struct CyclicScan {
    uint64_t *table[16];      // [rank]: that rank's table (local for rank == this rank)
    uint64_t *blockoff;       // local: 2 words per round {value, epoch}
    uint64_t *error;          // set when a wait for a peer times out
    uint64_t epoch;
    uint32_t rank, world, log2_block_tiles, table_stride; // entries per round in `table` (>= world)
};

struct ScanCall {
    cudaStream_t stream;
    const void *in;
    void *out;
    uint64_t size;
    uint64_t bs;       // == size for whole-array scans
    bool exclusive, reverse;
    const void *carry_in;
    void *carry_out;
    bool carry_api;    // whole array is one segment, optional carry
    const void *seeds = nullptr; // carry_api only: exclusive prefix of every 32 KiB tile (no look-back)
    const CyclicScan *cyclic = nullptr; // carry_api only: block-cyclic multi-GPU scan
};

/// Tries the streaming kernels (power-of-two blocks up to a tile; whole-array
/// and tile-aligned power-of-two blocks chained by look-back).  Sets *handled
/// to false when the call needs the general kernel (odd block sizes,
/// misaligned pointers).  'vt' / 'op' are already validated and canonicalised
/// like pick_scan() does.
int scan_fast_dispatch(int vt, int op, const ScanCall &call, bool *handled);

} // namespace b200
