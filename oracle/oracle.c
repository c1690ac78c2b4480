/*
 * oracle/oracle.c -- plain-C99 restatement of the reference algorithms for the
 * hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).  Every routine is the
 * obviously-correct serial form; the reference file:line it follows is cited
 * at each function.  All citations are relative to /root/reference.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* tests/reductions.cpp:5-13 */
uint32_t oracle_fmix32(uint32_t h) {
    h += 1;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

/* src/var.cpp:117-119 */
uint32_t oracle_type_size(int vt) {
    static const uint32_t ts[16] = { 0, 1, 0, 1, 1, 2, 2, 4, 4, 8, 8, 8, 0, 2, 4, 8 };
    return (vt >= 0 && vt < 16) ? ts[vt] : 0;
}

/* src/var.cpp:140-166 (type_all_ones / type_one / type_min / type_max) and
 * src/var.cpp:2642-2652 (jitc_reduce_identity) */
uint64_t oracle_reduce_identity(int vt, int op) {
    static const uint64_t all_ones[16] = {
        0, 1, 0, 0xff, 0xff, 0xffff, 0xffff, 0xffffffffu, 0xffffffffu,
        0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0, 0xffff, 0xffffffffu, 0xffffffffffffffffull };
    static const uint64_t one[16] = {
        0, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0x3c00, 0x3f800000,
        0x3ff0000000000000ull };
    static const uint64_t tmin[16] = {
        0, 0, 0, 0x80, 0, 0x8000, 0, 0x80000000u, 0, 0x8000000000000000ull, 0, 0,
        0, 0xfc00, 0xff800000u, 0xfff0000000000000ull };
    static const uint64_t tmax[16] = {
        0, 1, 0, 0x7f, 0xff, 0x7fff, 0xffff, 0x7fffffff, 0xffffffffu,
        0x7fffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull,
        0, 0x7c00, 0x7f800000, 0x7ff0000000000000ull };
    if (vt < 0 || vt >= 16)
        return 0;
    switch (op) {
        case ORACLE_OP_OR:
        case ORACLE_OP_ADD: return 0;
        case ORACLE_OP_AND: return all_ones[vt];
        case ORACLE_OP_MUL: return one[vt];
        case ORACLE_OP_MIN: return tmax[vt];
        case ORACLE_OP_MAX: return tmin[vt];
        default: return 0;
    }
}

/* ---- IEEE half <-> float (include/drjit-core/half.h provides the same
 * conversions in the reference; this is the textbook bit-level form) ---- */
float oracle_half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t) (127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t oracle_float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t) ((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) /* inf / nan */
        return (uint16_t) (sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0));
    if (exp >= 31)
        return (uint16_t) (sign | 0x7c00u); /* overflow -> inf */
    if (exp <= 0) {
        if (exp < -10)
            return (uint16_t) sign; /* underflow -> 0 */
        man |= 0x800000u;
        uint32_t shift = (uint32_t) (14 - exp);
        uint32_t hm = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1)))
            hm++;
        return (uint16_t) (sign | hm);
    }
    uint32_t hm = man >> 13, rem = man & 0x1fffu;
    uint16_t h = (uint16_t) (sign | ((uint32_t) exp << 10) | hm);
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1)))
        h++; /* may carry into the exponent, which is the correct rounding */
    return h;
}

/* ---- scalar reduction operators: src/llvm_red.h:10-80 (RedAdd..RedAnd),
 * float min/max as on the device (resources/common.h:66-91: fminf/fmaxf) ---- */
#define DEF_INT_OP(NAME, T)                                                    \
    static T NAME(int op, T a, T b) {                                          \
        switch (op) {                                                          \
            case ORACLE_OP_ADD: return (T) (a + b);                            \
            case ORACLE_OP_MUL: return (T) (a * b);                            \
            case ORACLE_OP_MIN: return a < b ? a : b;                          \
            case ORACLE_OP_MAX: return a > b ? a : b;                          \
            case ORACLE_OP_AND: return (T) (a & b);                            \
            case ORACLE_OP_OR:  return (T) (a | b);                            \
            default: return a;                                                 \
        }                                                                      \
    }

DEF_INT_OP(op_u8, uint8_t)
DEF_INT_OP(op_u32, uint32_t)
DEF_INT_OP(op_u64, uint64_t)

/* signed add/mul wrap like the unsigned kernels (src/cuda_ts.cpp:215-225) */
static int32_t op_i32(int op, int32_t a, int32_t b) {
    switch (op) {
        case ORACLE_OP_MIN: return a < b ? a : b;
        case ORACLE_OP_MAX: return a > b ? a : b;
        default: return (int32_t) op_u32(op, (uint32_t) a, (uint32_t) b);
    }
}
static int64_t op_i64(int op, int64_t a, int64_t b) {
    switch (op) {
        case ORACLE_OP_MIN: return a < b ? a : b;
        case ORACLE_OP_MAX: return a > b ? a : b;
        default: return (int64_t) op_u64(op, (uint64_t) a, (uint64_t) b);
    }
}
static float op_f32(int op, float a, float b) {
    switch (op) {
        case ORACLE_OP_ADD: return a + b;
        case ORACLE_OP_MUL: return a * b;
        case ORACLE_OP_MIN: return fminf(a, b);
        case ORACLE_OP_MAX: return fmaxf(a, b);
        default: return 0.f; /* src/llvm_red.h:53-80: bit ops on floats yield 0 */
    }
}
static double op_f64(int op, double a, double b) {
    switch (op) {
        case ORACLE_OP_ADD: return a + b;
        case ORACLE_OP_MUL: return a * b;
        case ORACLE_OP_MIN: return fmin(a, b);
        case ORACLE_OP_MAX: return fmax(a, b);
        default: return 0.0;
    }
}

static int op_valid(int vt, int op) {
    if (op < ORACLE_OP_ADD || op > ORACLE_OP_OR)
        return 0;
    switch (vt) {
        case ORACLE_VT_UINT8:
            return op == ORACLE_OP_AND || op == ORACLE_OP_OR;
        case ORACLE_VT_INT32: case ORACLE_VT_UINT32:
        case ORACLE_VT_INT64: case ORACLE_VT_UINT64:
            return 1;
        case ORACLE_VT_FLOAT16: case ORACLE_VT_FLOAT32: case ORACLE_VT_FLOAT64:
            return op != ORACLE_OP_AND && op != ORACLE_OP_OR;
        default:
            return 0;
    }
}

static double ident_f(int op) {
    switch (op) {
        case ORACLE_OP_MUL: return 1.0;
        case ORACLE_OP_MIN: return INFINITY;
        case ORACLE_OP_MAX: return -INFINITY;
        default: return 0.0;
    }
}

/* Generic element accessors so that the block loops below are written once.
 * 'Acc' is a tagged accumulator in the reference's Value type
 * (src/llvm_red.h:11: half accumulates as float) or in double when wide. */
typedef struct {
    int vt, op, wide;
    union { uint8_t u8; uint32_t u32; int32_t i32; uint64_t u64; int64_t i64;
            float f32; double f64; } v;
} Acc;

static void acc_init(Acc *a, int vt, int op, int wide) {
    a->vt = vt; a->op = op; a->wide = wide;
    uint64_t id = oracle_reduce_identity(vt, op);
    switch (vt) {
        case ORACLE_VT_UINT8:  a->v.u8 = (uint8_t) id; break;
        case ORACLE_VT_UINT32: a->v.u32 = (uint32_t) id; break;
        case ORACLE_VT_INT32:  a->v.i32 = (int32_t) (uint32_t) id; break;
        case ORACLE_VT_UINT64: a->v.u64 = id; break;
        case ORACLE_VT_INT64:  a->v.i64 = (int64_t) id; break;
        case ORACLE_VT_FLOAT16:
        case ORACLE_VT_FLOAT32:
            if (wide) a->v.f64 = ident_f(op); else a->v.f32 = (float) ident_f(op);
            break;
        case ORACLE_VT_FLOAT64: a->v.f64 = ident_f(op); break;
    }
}

static void acc_push(Acc *a, const void *in, size_t i) {
    switch (a->vt) {
        case ORACLE_VT_UINT8:  a->v.u8  = op_u8 (a->op, a->v.u8,  ((const uint8_t  *) in)[i]); break;
        case ORACLE_VT_UINT32: a->v.u32 = op_u32(a->op, a->v.u32, ((const uint32_t *) in)[i]); break;
        case ORACLE_VT_INT32:  a->v.i32 = op_i32(a->op, a->v.i32, ((const int32_t  *) in)[i]); break;
        case ORACLE_VT_UINT64: a->v.u64 = op_u64(a->op, a->v.u64, ((const uint64_t *) in)[i]); break;
        case ORACLE_VT_INT64:  a->v.i64 = op_i64(a->op, a->v.i64, ((const int64_t  *) in)[i]); break;
        case ORACLE_VT_FLOAT16: {
            float x = oracle_half_to_float(((const uint16_t *) in)[i]);
            if (a->wide) a->v.f64 = op_f64(a->op, a->v.f64, (double) x);
            else         a->v.f32 = op_f32(a->op, a->v.f32, x);
            break;
        }
        case ORACLE_VT_FLOAT32: {
            float x = ((const float *) in)[i];
            if (a->wide) a->v.f64 = op_f64(a->op, a->v.f64, (double) x);
            else         a->v.f32 = op_f32(a->op, a->v.f32, x);
            break;
        }
        case ORACLE_VT_FLOAT64: a->v.f64 = op_f64(a->op, a->v.f64, ((const double *) in)[i]); break;
    }
}

static void acc_store(const Acc *a, void *out, size_t i) {
    switch (a->vt) {
        case ORACLE_VT_UINT8:  ((uint8_t  *) out)[i] = a->v.u8; break;
        case ORACLE_VT_UINT32: ((uint32_t *) out)[i] = a->v.u32; break;
        case ORACLE_VT_INT32:  ((int32_t  *) out)[i] = a->v.i32; break;
        case ORACLE_VT_UINT64: ((uint64_t *) out)[i] = a->v.u64; break;
        case ORACLE_VT_INT64:  ((int64_t  *) out)[i] = a->v.i64; break;
        case ORACLE_VT_FLOAT16:
            ((uint16_t *) out)[i] = oracle_float_to_half(a->wide ? (float) a->v.f64 : a->v.f32);
            break;
        case ORACLE_VT_FLOAT32:
            ((float *) out)[i] = a->wide ? (float) a->v.f64 : a->v.f32;
            break;
        case ORACLE_VT_FLOAT64: ((double *) out)[i] = a->v.f64; break;
    }
}

/* Block reduction: src/llvm_red.h:87-122 with chunk_size == block_size (the
 * serial worker configuration, src/llvm_ts.cpp:300-303), equivalently
 * tests/reductions.cpp:15-36 (block_sum_ref).  Argument checks:
 * src/llvm_ts.cpp:269-278. */
int oracle_block_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                        const void *in, void *out, int wide) {
    if (size == 0)
        return 0;
    if (block_size == 0 || block_size > size)
        return 1;
    if (!op_valid(vt, op))
        return 2;
    uint32_t blocks = (size + block_size - 1) / block_size;
    for (uint32_t b = 0; b < blocks; ++b) {
        size_t start = (size_t) b * block_size, end = start + block_size;
        if (end > size)
            end = size;
        Acc a;
        acc_init(&a, vt, op, wide);
        for (size_t j = start; j < end; ++j)
            acc_push(&a, in, j);
        acc_store(&a, out, b);
    }
    return 0;
}

/* Block prefix reduction: src/llvm_red.h:131-179 with one chunk per block, i.e.
 * tests/reductions.cpp:48-70 (block_prefix_sum_ref) generalised to every op.
 * block_size == 1: src/llvm_ts.cpp:363-371.  In-place (in == out) is legal. */
int oracle_block_prefix_reduce(int vt, int op, uint32_t size, uint32_t block_size,
                               int exclusive, int reverse, const void *in,
                               void *out, int wide) {
    if (size == 0)
        return 0;
    if (block_size == 0 || block_size > size)
        return 1;
    if (!op_valid(vt, op) || vt == ORACLE_VT_UINT8)
        return 2;
    uint32_t blocks = (size + block_size - 1) / block_size;
    for (uint32_t b = 0; b < blocks; ++b) {
        size_t start = (size_t) b * block_size, end = start + block_size;
        if (end > size)
            end = size;
        Acc a, prev;
        acc_init(&a, vt, op, wide);
        if (!reverse) {
            for (size_t j = start; j < end; ++j) {
                prev = a;
                acc_push(&a, in, j);
                acc_store(exclusive ? &prev : &a, out, j);
            }
        } else {
            for (size_t j = end; j > start; --j) {
                prev = a;
                acc_push(&a, in, j - 1);
                acc_store(exclusive ? &prev : &a, out, j - 1);
            }
        }
    }
    return 0;
}

/* Dot product: src/llvm_red.h:226-238 -- a serial fma chain in the value type
 * (half: the reference's drjit::half fma rounds through float each step). */
int oracle_reduce_dot(int vt, const void *a, const void *b, uint32_t size,
                      void *out, int wide) {
    switch (vt) {
        case ORACLE_VT_FLOAT16: {
            const uint16_t *pa = (const uint16_t *) a, *pb = (const uint16_t *) b;
            if (wide) {
                double r = 0;
                for (uint32_t i = 0; i < size; ++i)
                    r += (double) oracle_half_to_float(pa[i]) * (double) oracle_half_to_float(pb[i]);
                *(uint16_t *) out = oracle_float_to_half((float) r);
            } else {
                uint16_t r = 0;
                for (uint32_t i = 0; i < size; ++i)
                    r = oracle_float_to_half(fmaf(oracle_half_to_float(pa[i]),
                                                  oracle_half_to_float(pb[i]),
                                                  oracle_half_to_float(r)));
                *(uint16_t *) out = r;
            }
            return 0;
        }
        case ORACLE_VT_FLOAT32: {
            const float *pa = (const float *) a, *pb = (const float *) b;
            if (wide) {
                double r = 0;
                for (uint32_t i = 0; i < size; ++i)
                    r = fma((double) pa[i], (double) pb[i], r);
                *(float *) out = (float) r;
            } else {
                float r = 0;
                for (uint32_t i = 0; i < size; ++i)
                    r = fmaf(pa[i], pb[i], r);
                *(float *) out = r;
            }
            return 0;
        }
        case ORACLE_VT_FLOAT64: {
            const double *pa = (const double *) a, *pb = (const double *) b;
            double r = 0;
            for (uint32_t i = 0; i < size; ++i)
                r = fma(pa[i], pb[i], r);
            *(double *) out = r;
            return 0;
        }
        default:
            return 2; /* src/llvm_red.h:247: unsupported data type */
    }
}

/* Mask compression: src/llvm_ts.cpp:706-780 (single work unit) */
uint32_t oracle_compress(const uint8_t *in, uint32_t size, uint32_t *out) {
    uint32_t accum = 0;
    for (uint32_t i = 0; i < size; ++i) {
        uint32_t value = (uint32_t) in[i];
        if (value)
            out[accum] = i;
        accum += value;
    }
    return accum;
}

/* Bucketing permutation: src/llvm_ts.cpp:785-933 with one task per group:
 * histogram (:838-861), bucket-major exclusive offsets + ascending-id offsets
 * records (:865-897), stable placement perm[offset[key]++] = i (:905-925). */
uint32_t oracle_block_mkperm(const uint32_t *values, uint32_t size,
                             uint32_t block_size, uint32_t bucket_count,
                             uint32_t *perm, uint32_t *offsets) {
    if (size == 0 || bucket_count == 0 || block_size == 0)
        return 0;
    uint32_t n_groups = (size + block_size - 1) / block_size;
    uint32_t *hist = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) bucket_count);
    uint32_t unique = 0;
    if (!hist)
        return UINT32_MAX;
    for (uint32_t g = 0; g < n_groups; ++g) {
        size_t start = (size_t) g * block_size, end = start + block_size;
        if (end > size)
            end = size;
        memset(hist, 0, sizeof(uint32_t) * (size_t) bucket_count);
        for (size_t i = start; i < end; ++i) {
            if (values[i] >= bucket_count) { free(hist); return UINT32_MAX; }
            hist[values[i]]++;
        }
        uint32_t group_offset = 0;
        for (uint32_t b = 0; b < bucket_count; ++b) {
            uint32_t count = hist[b];
            hist[b] = (uint32_t) start + group_offset;
            if (n_groups == 1 && count > 0 && offsets) {
                offsets[unique * 4 + 0] = b;
                offsets[unique * 4 + 1] = group_offset;
                offsets[unique * 4 + 2] = count;
                offsets[unique * 4 + 3] = 0;
                unique++;
            }
            group_offset += count;
        }
        for (size_t i = start; i < end; ++i)
            perm[hist[values[i]]++] = (uint32_t) i;
    }
    free(hist);
    if (offsets && n_groups == 1) {
        offsets[4 * (size_t) bucket_count] = unique;
        return unique;
    }
    return 0;
}

/* Scatter-reduce: the operation every variant of src/cuda_scatter.cpp:246-354
 * implements (Direct / Local / NoConflicts differ only in how conflicts are
 * combined), applied serially.  Legal pairs: src/op.cpp:2735-2820 -- no Mul, no
 * And/Or on floats, no 8-bit types.  Float min/max follow the integer-atomic
 * emulation (src/cuda_scatter.cpp:74-106), which orders by bit pattern and so
 * equals fmin/fmax on non-NaN data. */
int oracle_scatter_reduce(int vt, int op, void *target, uint32_t target_size,
                          const void *value, const uint32_t *index,
                          const uint8_t *mask, uint32_t n, int wide) {
    if (op == ORACLE_OP_MUL || op == ORACLE_OP_IDENTITY || !op_valid(vt, op) ||
        vt == ORACLE_VT_UINT8)
        return 2;
    int is_f = vt == ORACLE_VT_FLOAT16 || vt == ORACLE_VT_FLOAT32;
    double *shadow = NULL;
    if (wide && is_f && op == ORACLE_OP_ADD) {
        shadow = (double *) malloc(sizeof(double) * (size_t) target_size);
        if (!shadow)
            return 3;
        for (uint32_t i = 0; i < target_size; ++i)
            shadow[i] = vt == ORACLE_VT_FLOAT32
                            ? (double) ((float *) target)[i]
                            : (double) oracle_half_to_float(((uint16_t *) target)[i]);
    }
    for (uint32_t i = 0; i < n; ++i) {
        if (mask && !mask[i])
            continue;
        uint32_t j = index[i];
        if (j >= target_size) { free(shadow); return 4; }
        switch (vt) {
            case ORACLE_VT_UINT32: ((uint32_t *) target)[j] = op_u32(op, ((uint32_t *) target)[j], ((const uint32_t *) value)[i]); break;
            case ORACLE_VT_INT32:  ((int32_t  *) target)[j] = op_i32(op, ((int32_t  *) target)[j], ((const int32_t  *) value)[i]); break;
            case ORACLE_VT_UINT64: ((uint64_t *) target)[j] = op_u64(op, ((uint64_t *) target)[j], ((const uint64_t *) value)[i]); break;
            case ORACLE_VT_INT64:  ((int64_t  *) target)[j] = op_i64(op, ((int64_t  *) target)[j], ((const int64_t  *) value)[i]); break;
            case ORACLE_VT_FLOAT64: ((double *) target)[j] = op_f64(op, ((double *) target)[j], ((const double *) value)[i]); break;
            case ORACLE_VT_FLOAT32:
                if (shadow) shadow[j] += (double) ((const float *) value)[i];
                else ((float *) target)[j] = op_f32(op, ((float *) target)[j], ((const float *) value)[i]);
                break;
            case ORACLE_VT_FLOAT16: {
                float x = oracle_half_to_float(((const uint16_t *) value)[i]);
                if (shadow) shadow[j] += (double) x;
                else ((uint16_t *) target)[j] = oracle_float_to_half(
                         op_f32(op, oracle_half_to_float(((uint16_t *) target)[j]), x));
                break;
            }
        }
    }
    if (shadow) {
        for (uint32_t i = 0; i < target_size; ++i) {
            if (vt == ORACLE_VT_FLOAT32) ((float *) target)[i] = (float) shadow[i];
            else ((uint16_t *) target)[i] = oracle_float_to_half((float) shadow[i]);
        }
        free(shadow);
    }
    return 0;
}

/* src/util.cpp:172-211: all()/any() */
int oracle_all(const uint8_t *values, uint32_t size) {
    for (uint32_t i = 0; i < size; ++i)
        if (!values[i])
            return 0;
    return 1;
}

int oracle_any(const uint8_t *values, uint32_t size) {
    for (uint32_t i = 0; i < size; ++i)
        if (values[i])
            return 1;
    return 0;
}
