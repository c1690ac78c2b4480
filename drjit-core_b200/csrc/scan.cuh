// scan.cuh -- interface between the prefix-reduction dispatcher (scan.cu, which
// also holds the fully general segmented kernel) and the streaming fast paths
// (scan_fast.cu).
#pragma once

#include "common.cuh"

namespace b200 {

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
};

/// Tries the streaming kernels (power-of-two blocks up to a tile; whole-array
/// and tile-aligned power-of-two blocks chained by look-back).  Sets *handled
/// to false when the call needs the general kernel (odd block sizes,
/// misaligned pointers).  'vt' / 'op' are already validated and canonicalised
/// like pick_scan() does.
int scan_fast_dispatch(int vt, int op, const ScanCall &call, bool *handled);

} // namespace b200
