/*
 * hj_oracle.c — CPU restatement of the reference's device-op algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may load it,
 * and only as the checker or the timed CPU baseline.  The product path (the CUDA library in
 * hephaestus-jit_b200/) never links, imports or falls back to this file.
 *
 * Why a restatement and not the reference itself: the reference is nightly Rust + Vulkan
 * (GLSL -> SPIR-V through shaderc, executed by a Vulkan ICD).  Neither cargo/rustc nor a
 * Vulkan loader/ICD/shaderc exist in this image, so it cannot be built or
 * run here.  Every function below cites the reference lines it follows (paths relative to
 * hephaestus-jit/src/backend/vulkan/builtin/).  Pinning: tests/test_oracle_golden.py checks
 * this file against every known-answer vector the reference's own tests hold for the path
 * (hephaestus-jit/src/test.rs:493-1019, transcribed in tests/golden/reference_kats.json).
 * PARITY UNPINNED where the reference holds no vector: exclusive scans (the
 * reference's exclusive scan is inclusive), f32 / f64 / u64 scans (D3), reductions
 * beyond 1000 elements, compress with a non-trivial mask beyond 128 elements, and every float
 * elementwise op other than cos.  For those this file is the only statement of the contract.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off, no -march=native so the
 * .so also runs on the GPU box's host CPU; contraction is off so float results do not
 * depend on whether the host has FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hj.h"

#define HJO_OK 0
#define HJO_INVALID (-1)
#define HJO_UNSUPPORTED (-2)
#define HJO_OOM (-3)

static int g_threads = 0; /* 0 = OpenMP default (all cores) */

void hjo_set_threads(int n) { g_threads = n; }
int hjo_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
#ifdef _OPENMP
#define NTHREADS() (g_threads > 0 ? g_threads : omp_get_max_threads())
#else
#define NTHREADS() 1
#endif

size_t hjo_type_size(int ty) {
    switch (ty) {
    case HJ_BOOL: case HJ_I8: case HJ_U8: return 1;
    case HJ_I16: case HJ_U16: case HJ_F16: return 2;
    case HJ_I32: case HJ_U32: case HJ_F32: return 4;
    case HJ_I64: case HJ_U64: case HJ_F64: return 8;
    default: return 0;
    }
}

/* =====================================================================================
 * Reduction — follows builtin/reduce.rs:22-314 and kernels/reduce.glsl:1-50.
 *
 *   n_passes     = (num - 1).ilog(32) + 1                       reduce.rs:33
 *   scratch_size = 32^n_passes                                  reduce.rs:34
 *   pass i (i = n_passes-1 .. 0) launches 32^i workgroups of 32 reduce.rs:245-290
 *   each workgroup:  sh[l] = gid < size ? in[gid] : INIT        reduce.glsl:35
 *                    for s = 16,8,4,2,1: if l < s: sh[l] = REDUCE(sh[l], sh[l+s])
 *                                                               reduce.glsl:40-47
 *                    out[group] = sh[0]                         reduce.glsl:49
 *   INIT per (op, type): reduce.rs:84-166;  REDUCE per op: reduce.rs:74-83.
 *
 * The staging copy src -> scratch (reduce.rs:219-240) is the identity on values, so the
 * first level reads `src` directly unless `faithful_copy` asks for the memcpy (used when
 * this function is timed as the CPU baseline so the work matches the reference's).
 * Note `size` stays the ORIGINAL num in every pass (reduce.rs:70-72 fills the size buffer
 * once); at level k>0 all 32^(n_passes-k) inputs are < num, so the test never fires there.
 * Deviation D6: the reference panics for num == 1 (ilog(0)); we return
 * src[0].  num == 0 is rejected.
 * ===================================================================================== */

/* GLSL min/max (GLSL 4.60 §8.3): min(x,y) = y < x ? y : x ; max(x,y) = x < y ? y : x */
#define OP_MAX(a, b) ((a) < (b) ? (b) : (a))
#define OP_MIN(a, b) ((b) < (a) ? (b) : (a))

#define DEF_TREE(NAME, T, EXPR)                                                           \
    static void NAME(void* p) {                                                           \
        T* sh = (T*)p;                                                                    \
        for (unsigned s = 16; s > 0; s >>= 1)                                             \
            for (unsigned l = 0; l < s; l++) {                                            \
                T a = sh[l], b = sh[l + s];                                               \
                sh[l] = (T)(EXPR);                                                        \
            }                                                                             \
    }

/* integer arithmetic is done in the unsigned type of the same width: wrapping, no UB */
#define DEF_INT_TREES(SFX, T, UT)                                                         \
    DEF_TREE(tree_max_##SFX, T, OP_MAX(a, b))                                             \
    DEF_TREE(tree_min_##SFX, T, OP_MIN(a, b))                                             \
    DEF_TREE(tree_sum_##SFX, T, (UT)((UT)a + (UT)b))                                      \
    DEF_TREE(tree_prod_##SFX, T, (UT)((UT)a * (UT)b))                                     \
    DEF_TREE(tree_or_##SFX, T, (UT)((UT)a | (UT)b))                                       \
    DEF_TREE(tree_and_##SFX, T, (UT)((UT)a & (UT)b))                                      \
    DEF_TREE(tree_xor_##SFX, T, (UT)((UT)a ^ (UT)b))

DEF_INT_TREES(i8, int8_t, uint8_t)
DEF_INT_TREES(u8, uint8_t, uint8_t)
DEF_INT_TREES(i16, int16_t, uint16_t)
DEF_INT_TREES(u16, uint16_t, uint16_t)
DEF_INT_TREES(i32, int32_t, uint32_t)
DEF_INT_TREES(u32, uint32_t, uint32_t)
DEF_INT_TREES(i64, int64_t, uint64_t)
DEF_INT_TREES(u64, uint64_t, uint64_t)

#define DEF_FLOAT_TREES(SFX, T)                                                           \
    DEF_TREE(tree_max_##SFX, T, OP_MAX(a, b))                                             \
    DEF_TREE(tree_min_##SFX, T, OP_MIN(a, b))                                             \
    DEF_TREE(tree_sum_##SFX, T, a + b)                                                    \
    DEF_TREE(tree_prod_##SFX, T, a * b)

DEF_FLOAT_TREES(f32, float)
DEF_FLOAT_TREES(f64, double)

typedef void (*tree32_fn)(void*);

typedef struct {
    tree32_fn fn;
    size_t es;
    uint64_t init; /* bit pattern of INIT, low `es` bytes */
} reduce_desc;

static uint64_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t f64_bits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }

/* INIT table: reduce.rs:84-166.  Anything the reference leaves as todo!() -> unsupported. */
static int reduce_lookup(int op, int ty, reduce_desc* d) {
#define INT_CASES(SFX, ES, MAXINIT, MININIT, ALLONES)                                     \
    d->es = ES;                                                                           \
    switch (op) {                                                                         \
    case HJ_REDUCE_MAX: d->fn = tree_max_##SFX; d->init = (uint64_t)(MAXINIT); return 0;  \
    case HJ_REDUCE_MIN: d->fn = tree_min_##SFX; d->init = (uint64_t)(MININIT); return 0;  \
    case HJ_REDUCE_SUM: d->fn = tree_sum_##SFX; d->init = 0; return 0;                    \
    case HJ_REDUCE_PROD: d->fn = tree_prod_##SFX; d->init = 1; return 0;                  \
    default: break;                                                                       \
    }
    switch (ty) {
    case HJ_BOOL: /* reduce.rs:141,150,159: Bool only for And/Or/Xor, stored as u8 */
        d->es = 1;
        if (op == HJ_REDUCE_AND) { d->fn = tree_and_u8; d->init = 1; return 0; }
        if (op == HJ_REDUCE_OR) { d->fn = tree_or_u8; d->init = 0; return 0; }
        if (op == HJ_REDUCE_XOR) { d->fn = tree_xor_u8; d->init = 0; return 0; }
        return HJO_UNSUPPORTED;
    case HJ_I8: INT_CASES(i8, 1, 0x80u, 0x7fu, 0) return HJO_UNSUPPORTED;
    case HJ_U8:
        INT_CASES(u8, 1, 0u, 0xffu, 0)
        if (op == HJ_REDUCE_AND) { d->fn = tree_and_u8; d->init = 0xffu; return 0; }
        if (op == HJ_REDUCE_OR) { d->fn = tree_or_u8; d->init = 0; return 0; }
        if (op == HJ_REDUCE_XOR) { d->fn = tree_xor_u8; d->init = 0; return 0; }
        return HJO_UNSUPPORTED;
    case HJ_I16: INT_CASES(i16, 2, 0x8000u, 0x7fffu, 0) return HJO_UNSUPPORTED;
    case HJ_U16:
        INT_CASES(u16, 2, 0u, 0xffffu, 0)
        if (op == HJ_REDUCE_AND) { d->fn = tree_and_u16; d->init = 0xffffu; return 0; }
        if (op == HJ_REDUCE_OR) { d->fn = tree_or_u16; d->init = 0; return 0; }
        if (op == HJ_REDUCE_XOR) { d->fn = tree_xor_u16; d->init = 0; return 0; }
        return HJO_UNSUPPORTED;
    case HJ_I32: INT_CASES(i32, 4, 0x80000000u, 0x7fffffffu, 0) return HJO_UNSUPPORTED;
    case HJ_U32:
        INT_CASES(u32, 4, 0u, 0xffffffffu, 0)
        if (op == HJ_REDUCE_AND) { d->fn = tree_and_u32; d->init = 0xffffffffu; return 0; }
        if (op == HJ_REDUCE_OR) { d->fn = tree_or_u32; d->init = 0; return 0; }
        if (op == HJ_REDUCE_XOR) { d->fn = tree_xor_u32; d->init = 0; return 0; }
        return HJO_UNSUPPORTED;
    case HJ_I64:
        INT_CASES(i64, 8, 0x8000000000000000ull, 0x7fffffffffffffffull, 0)
        return HJO_UNSUPPORTED;
    case HJ_U64:
        INT_CASES(u64, 8, 0ull, 0xffffffffffffffffull, 0)
        if (op == HJ_REDUCE_AND) { d->fn = tree_and_u64; d->init = ~0ull; return 0; }
        if (op == HJ_REDUCE_OR) { d->fn = tree_or_u64; d->init = 0; return 0; }
        if (op == HJ_REDUCE_XOR) { d->fn = tree_xor_u64; d->init = 0; return 0; }
        return HJO_UNSUPPORTED;
    case HJ_F32:
        d->es = 4;
        switch (op) {
        case HJ_REDUCE_MAX: d->fn = tree_max_f32; d->init = f32_bits(-INFINITY); return 0;
        case HJ_REDUCE_MIN: d->fn = tree_min_f32; d->init = f32_bits(INFINITY); return 0;
        case HJ_REDUCE_SUM: d->fn = tree_sum_f32; d->init = f32_bits(0.f); return 0;
        case HJ_REDUCE_PROD: d->fn = tree_prod_f32; d->init = f32_bits(1.f); return 0;
        default: return HJO_UNSUPPORTED;
        }
    case HJ_F64:
        d->es = 8;
        switch (op) {
        case HJ_REDUCE_MAX: d->fn = tree_max_f64; d->init = f64_bits(-INFINITY); return 0;
        case HJ_REDUCE_MIN: d->fn = tree_min_f64; d->init = f64_bits(INFINITY); return 0;
        case HJ_REDUCE_SUM: d->fn = tree_sum_f64; d->init = f64_bits(0.0); return 0;
        case HJ_REDUCE_PROD: d->fn = tree_prod_f64; d->init = f64_bits(1.0); return 0;
        default: return HJO_UNSUPPORTED;
        }
    default: /* F16 and Void: todo!() in reduce.rs */
        return HJO_UNSUPPORTED;
    }
#undef INT_CASES
}

/* 1 if the reference supports (op, ty) — lets the tests enumerate the table. */
int hjo_reduce_supported(int op, int ty) {
    reduce_desc d;
    return reduce_lookup(op, ty, &d) == 0;
}

static void fill_init(void* dst, size_t count, size_t es, uint64_t init) {
    unsigned char* p = (unsigned char*)dst;
    for (size_t i = 0; i < count; i++) memcpy(p + i * es, &init, es);
}

/* one pass: out[g] = tree32(in[32g .. 32g+32)) with elements >= n_valid replaced by INIT */
static void reduce_level(const reduce_desc* d, const unsigned char* in, size_t n_valid,
                         unsigned char* out, size_t n_groups) {
    const size_t es = d->es;
#pragma omp parallel for schedule(static) num_threads(NTHREADS())
    for (long long g = 0; g < (long long)n_groups; g++) {
        uint64_t shbuf[32]; /* 32 elements of up to 8 bytes, 8-byte aligned */
        unsigned char* sh = (unsigned char*)shbuf;
        size_t base = (size_t)g * 32;
        if (base >= n_valid) { /* all-INIT group: REDUCE(INIT, INIT) == INIT for every row */
            memcpy(out + (size_t)g * es, &d->init, es);
            continue;
        }
        size_t have = n_valid - base < 32 ? n_valid - base : 32;
        memcpy(sh, in + base * es, have * es);
        if (have < 32) fill_init(sh + have * es, 32 - have, es, d->init);
        d->fn(sh);
        memcpy(out + (size_t)g * es, sh, es);
    }
}

int hjo_reduce(int op, int ty, size_t n, const void* src, void* dst, int faithful_copy) {
    reduce_desc d;
    int rc = reduce_lookup(op, ty, &d);
    if (rc) return rc;
    if (n == 0 || !src || !dst) return HJO_INVALID;
    if (n == 1) { memcpy(dst, src, d.es); return HJO_OK; } /* D6 */

    /* n_passes = ilog32(n-1) + 1 */
    unsigned n_passes = 0;
    for (size_t v = n - 1; v > 0; v /= 32) n_passes++;
    if (n_passes == 0) n_passes = 1;
    size_t groups = 1;
    for (unsigned i = 1; i < n_passes; i++) groups *= 32; /* 32^(n_passes-1) */

    const unsigned char* in = (const unsigned char*)src;
    unsigned char* staged = NULL;
    if (faithful_copy) { /* reduce.rs:219-240: vkCmdCopyBuffer(src -> scratch) */
        staged = (unsigned char*)malloc(n * d.es);
        if (!staged) return HJO_OOM;
#pragma omp parallel for schedule(static) num_threads(NTHREADS())
        for (long long c = 0; c < (long long)((n * d.es + (1 << 20) - 1) >> 20); c++) {
            size_t off = (size_t)c << 20, len = n * d.es - off;
            if (len > (1u << 20)) len = 1u << 20;
            memcpy(staged + off, in + off, len);
        }
        in = staged;
    }
    unsigned char* a = (unsigned char*)malloc(groups * d.es);
    unsigned char* b = (unsigned char*)malloc((groups / 32 + 1) * d.es);
    if (!a || !b) { free(a); free(b); free(staged); return HJO_OOM; }

    size_t n_valid = n;
    unsigned char* out = a;
    unsigned char* other = b;
    for (unsigned pass = 0; pass < n_passes; pass++) {
        reduce_level(&d, in, n_valid, out, groups);
        in = out;
        n_valid = groups; /* every produced element is "< size" (see header comment) */
        groups /= 32;
        unsigned char* t = out; out = other; other = t;
    }
    memcpy(dst, in, d.es); /* reduce.rs:292-313: copy one element to dst */
    free(a); free(b); free(staged);
    return HJO_OK;
}

/* =====================================================================================
 * Prefix sum — follows builtin/prefix_sum.rs:31-162 and kernels/prefix_sum_large.glsl.
 *
 *   block_size 128, N = 4 loads of M = 4-wide vectors -> 16 items/thread,
 *   2048 items per partition                                   prefix_sum.rs:40-47
 *   out-of-range items are zeroed on load (INIT)               prefix_sum_large.glsl:190-207
 *   shared-memory transpose: thread t owns items [16t, 16t+16) prefix_sum_large.glsl:209-226
 *   serial scan of the 16 items (inclusive or exclusive)       prefix_sum_large.glsl:229-241
 *   Hillis-Steele inclusive scan of the 128 thread sums        prefix_sum_large.glsl:246-259
 *   look-back: prefix = sum of predecessors' aggregates        prefix_sum_large.glsl:281-317
 *   values[i] += (sum_block + prefix) - sum_local              prefix_sum_large.glsl:329-332
 *
 * The look-back's summation ORDER depends on timing in the reference (which predecessor
 * already published flag 2); we restate the fully-serialised schedule, in which partition
 * p sees partition p-1 complete: prefix_p = inclusive total through p-1.  For integers
 * every order gives the same wrapped result, so integer scans are bit-exact by
 * construction; for floats this fixes one of the reference's possible answers.
 *
 * Deviations: D3 — the reference packs every type's prefix with the u32
 * rule, corrupting f32/u64/f64 carries; we carry the value in its own type.  D10 — the
 * reference's `inclusive=false` still produces an inclusive scan; `inclusive` here is
 * honoured, and `ref_compat != 0` reproduces the always-inclusive behaviour.
 * Writes exactly n outputs (the reference over-writes up to the partition end, D7).
 * ===================================================================================== */

#define SCAN_BLOCK 128
#define SCAN_ITEMS 16
#define SCAN_PART (SCAN_BLOCK * SCAN_ITEMS)

#define DEF_SCAN(SFX, T, AT)                                                              \
    static void scan_##SFX(const T* src, T* dst, size_t n, int inclusive) {               \
        size_t n_parts = (n + SCAN_PART - 1) / SCAN_PART;                                 \
        AT carry = 0; /* inclusive total through the previous partition */                \
        T* part = (T*)malloc(sizeof(T) * SCAN_PART);                                      \
        for (size_t p = 0; p < n_parts; p++) {                                            \
            size_t base = p * SCAN_PART;                                                  \
            AT sum_local[SCAN_BLOCK];                                                     \
            for (unsigned t = 0; t < SCAN_BLOCK; t++) {                                   \
                AT s = 0;                                                                 \
                for (unsigned i = 0; i < SCAN_ITEMS; i++) {                               \
                    size_t j = base + (size_t)t * SCAN_ITEMS + i;                         \
                    AT v = j < n ? (AT)src[j] : (AT)0;                                    \
                    if (inclusive) { s = (AT)(s + v); part[t * SCAN_ITEMS + i] = (T)s; }  \
                    else { part[t * SCAN_ITEMS + i] = (T)s; s = (AT)(s + v); }            \
                }                                                                         \
                sum_local[t] = s;                                                         \
            }                                                                             \
            /* Hillis-Steele, offsets 1,2,4,...,64 with a zero guard band */              \
            AT hs[SCAN_BLOCK], tmp[SCAN_BLOCK];                                           \
            memcpy(hs, sum_local, sizeof(hs));                                            \
            for (unsigned off = 1; off < SCAN_BLOCK; off <<= 1) {                         \
                for (unsigned t = 0; t < SCAN_BLOCK; t++)                                 \
                    tmp[t] = (AT)(hs[t] + (t >= off ? hs[t - off] : (AT)0));              \
                memcpy(hs, tmp, sizeof(hs));                                              \
            }                                                                             \
            AT prefix = carry;                                                            \
            for (unsigned t = 0; t < SCAN_BLOCK; t++) {                                   \
                AT sum_block = (AT)(hs[t] + prefix);                                      \
                AT add = (AT)(sum_block - sum_local[t]);                                  \
                for (unsigned i = 0; i < SCAN_ITEMS; i++) {                               \
                    size_t j = base + (size_t)t * SCAN_ITEMS + i;                         \
                    if (j < n) dst[j] = (T)((AT)part[t * SCAN_ITEMS + i] + add);          \
                }                                                                         \
            }                                                                             \
            carry = (AT)(hs[SCAN_BLOCK - 1] + prefix);                                    \
        }                                                                                 \
        free(part);                                                                       \
    }

DEF_SCAN(u8, uint8_t, uint8_t)
DEF_SCAN(i8, int8_t, uint8_t)
DEF_SCAN(u16, uint16_t, uint16_t)
DEF_SCAN(i16, int16_t, uint16_t)
DEF_SCAN(u32, uint32_t, uint32_t)
DEF_SCAN(i32, int32_t, uint32_t)
DEF_SCAN(u64, uint64_t, uint64_t)
DEF_SCAN(i64, int64_t, uint64_t)
DEF_SCAN(f32, float, float)
DEF_SCAN(f64, double, double)

int hjo_prefix_sum(int ty, size_t n, int inclusive, int ref_compat, const void* src, void* dst) {
    if (n == 0 || !src || !dst) return HJO_INVALID;
    if (ref_compat) inclusive = 1; /* D10 */
    switch (ty) {
    case HJ_U8: scan_u8(src, dst, n, inclusive); return 0;
    case HJ_I8: scan_i8(src, dst, n, inclusive); return 0;
    case HJ_U16: scan_u16(src, dst, n, inclusive); return 0;
    case HJ_I16: scan_i16(src, dst, n, inclusive); return 0;
    case HJ_U32: scan_u32(src, dst, n, inclusive); return 0;
    case HJ_I32: scan_i32(src, dst, n, inclusive); return 0;
    case HJ_U64: scan_u64(src, dst, n, inclusive); return 0;
    case HJ_I64: scan_i64(src, dst, n, inclusive); return 0;
    case HJ_F32: scan_f32(src, dst, n, inclusive); return 0;
    case HJ_F64: scan_f64(src, dst, n, inclusive); return 0;
    default: return HJO_UNSUPPORTED;
    }
}

/* Multi-threaded scan with the SAME results for integer types (three-phase: per-chunk
 * totals, serial scan of totals, per-chunk scan seeded with the offset).  Used only as the
 * timed CPU baseline for integer scans; chunks are multiples of the reference partition. */
int hjo_prefix_sum_u32_mt(size_t n, int inclusive, const uint32_t* src, uint32_t* dst) {
    if (n == 0 || !src || !dst) return HJO_INVALID;
    int nt = NTHREADS();
    size_t chunk = ((n + nt - 1) / nt + SCAN_PART - 1) / SCAN_PART * SCAN_PART;
    size_t n_chunks = (n + chunk - 1) / chunk;
    uint32_t* totals = (uint32_t*)calloc(n_chunks + 1, sizeof(uint32_t));
    if (!totals) return HJO_OOM;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long c = 0; c < (long long)n_chunks; c++) {
        size_t b = (size_t)c * chunk, e = b + chunk < n ? b + chunk : n;
        uint32_t s = 0;
        for (size_t i = b; i < e; i++) s += src[i];
        totals[c + 1] = s;
    }
    for (size_t c = 0; c < n_chunks; c++) totals[c + 1] += totals[c];
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long c = 0; c < (long long)n_chunks; c++) {
        size_t b = (size_t)c * chunk, e = b + chunk < n ? b + chunk : n;
        uint32_t s = totals[c];
        if (inclusive) for (size_t i = b; i < e; i++) { s += src[i]; dst[i] = s; }
        else for (size_t i = b; i < e; i++) { uint32_t v = src[i]; dst[i] = s; s += v; }
    }
    free(totals);
    return HJO_OK;
}

/* =====================================================================================
 * Compress — follows builtin/compress.rs:157-283 and kernels/compress_large.glsl:76-229.
 *
 *   block_size 128, 16 mask bytes per thread (one u32vec4)     compress.rs:166-172,
 *                                                              compress_large.glsl:101-123
 *   serial EXCLUSIVE scan of the 16 byte values + a 17th zero  compress_large.glsl:126-132
 *   Hillis-Steele block scan, look-back prefix                 compress_large.glsl:134-196
 *   last partition's last thread writes out_count              compress_large.glsl:214-216
 *   for i in 0..16: if values[i] != values[i+1]:
 *        out[values[i]] = (partition*128 + thread)*16 + i      compress_large.glsl:224-228
 *
 * The rank is the running sum of the mask BYTE VALUES; the mask is a `bool` buffer whose
 * memory type is u8 holding 0/1 (codegen/glsl/mod.rs:256-258), so rank == number of
 * preceding set elements.  Callers must pass 0/1 bytes.
 * Deviation D4: the reference never masks elements >= size and relies on
 * zeroed padding; we treat bytes at and beyond n as 0.
 * index_out entries at and beyond the count are NOT written (they keep the zero fill of the
 * preceding literal kernel, hephaestus-jit/src/trace.rs:1600-1601).
 * ===================================================================================== */
#define CMP_BLOCK 128
#define CMP_ITEMS 16
#define CMP_PART (CMP_BLOCK * CMP_ITEMS)

int hjo_compress(size_t n, const uint8_t* mask, uint32_t* index_out, uint32_t* out_count,
                 uint32_t index_base) {
    if (n == 0 || !mask || !index_out || !out_count) return HJO_INVALID;
    size_t n_parts = (n + CMP_PART - 1) / CMP_PART;
    uint32_t carry = 0;
    for (size_t p = 0; p < n_parts; p++) {
        uint32_t values[CMP_BLOCK][CMP_ITEMS + 1];
        uint32_t sum_local[CMP_BLOCK], hs[CMP_BLOCK], tmp[CMP_BLOCK];
        for (unsigned t = 0; t < CMP_BLOCK; t++) {
            uint32_t s = 0;
            for (unsigned i = 0; i < CMP_ITEMS + 1; i++) {
                size_t j = (p * CMP_BLOCK + t) * CMP_ITEMS + i;
                uint32_t v = (i < CMP_ITEMS && j < n) ? mask[j] : 0u;
                values[t][i] = s;
                s += v;
            }
            sum_local[t] = s;
        }
        memcpy(hs, sum_local, sizeof(hs));
        for (unsigned off = 1; off < CMP_BLOCK; off <<= 1) {
            for (unsigned t = 0; t < CMP_BLOCK; t++) tmp[t] = hs[t] + (t >= off ? hs[t - off] : 0u);
            memcpy(hs, tmp, sizeof(hs));
        }
        for (unsigned t = 0; t < CMP_BLOCK; t++) {
            uint32_t add = hs[t] + carry - sum_local[t];
            for (unsigned i = 0; i < CMP_ITEMS; i++) {
                uint32_t lo = values[t][i] + add, hi = values[t][i + 1] + add;
                if (lo != hi)
                    index_out[lo] =
                        (uint32_t)((p * CMP_BLOCK + t) * CMP_ITEMS + i) + index_base;
            }
        }
        carry += hs[CMP_BLOCK - 1];
    }
    *out_count = carry;
    return HJO_OK;
}

/* Multi-threaded compress with identical output (count pass, serial offsets, emit pass);
 * timed CPU baseline only. */
int hjo_compress_mt(size_t n, const uint8_t* mask, uint32_t* index_out, uint32_t* out_count,
                    uint32_t index_base) {
    if (n == 0 || !mask || !index_out || !out_count) return HJO_INVALID;
    int nt = NTHREADS();
    size_t chunk = ((n + nt - 1) / nt + CMP_PART - 1) / CMP_PART * CMP_PART;
    size_t n_chunks = (n + chunk - 1) / chunk;
    uint32_t* totals = (uint32_t*)calloc(n_chunks + 1, sizeof(uint32_t));
    if (!totals) return HJO_OOM;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long c = 0; c < (long long)n_chunks; c++) {
        size_t b = (size_t)c * chunk, e = b + chunk < n ? b + chunk : n;
        uint32_t s = 0;
        for (size_t i = b; i < e; i++) s += mask[i];
        totals[c + 1] = s;
    }
    for (size_t c = 0; c < n_chunks; c++) totals[c + 1] += totals[c];
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long long c = 0; c < (long long)n_chunks; c++) {
        size_t b = (size_t)c * chunk, e = b + chunk < n ? b + chunk : n;
        uint32_t r = totals[c];
        for (size_t i = b; i < e; i++)
            if (mask[i]) index_out[r++] = (uint32_t)i + index_base;
    }
    *out_count = totals[n_chunks];
    free(totals);
    return HJO_OK;
}

/* =====================================================================================
 * ScatterReduce / Gather kernel ops — semantics of codegen/glsl/mod.rs:400-446 (atomicOp on
 * buffer[idx]) and :523-579 (buffer[idx] load), applied for i = 0..n in index order.
 * Integer results do not depend on the order; f32 sums do (tolerance in the tests).
 * ===================================================================================== */
int hjo_scatter_reduce(int op, int ty, size_t n, const uint32_t* idx, const void* src,
                       uint64_t literal, void* dst, size_t n_dst) {
    if (!idx || !dst) return HJO_INVALID;
    if (op == HJ_REDUCE_PROD) return HJO_UNSUPPORTED; /* todo!() at glsl/mod.rs:422 */
#define SR_LOOP(T, UT, ISINT)                                                             \
    {                                                                                     \
        T* d = (T*)dst;                                                                   \
        const T* s = (const T*)src;                                                       \
        T lit; memcpy(&lit, &literal, sizeof(T));                                         \
        for (size_t i = 0; i < n; i++) {                                                  \
            uint32_t k = idx[i];                                                          \
            if (k >= n_dst) return HJO_INVALID;                                           \
            T v = s ? s[i] : lit, a = d[k];                                               \
            switch (op) {                                                                 \
            case HJ_REDUCE_MAX: d[k] = OP_MAX(a, v); break;                               \
            case HJ_REDUCE_MIN: d[k] = OP_MIN(a, v); break;                               \
            case HJ_REDUCE_SUM: d[k] = ISINT ? (T)((UT)a + (UT)v) : (T)(a + v); break;    \
            case HJ_REDUCE_OR: if (ISINT) d[k] = (T)((UT)a | (UT)v); break;               \
            case HJ_REDUCE_AND: if (ISINT) d[k] = (T)((UT)a & (UT)v); break;              \
            case HJ_REDUCE_XOR: if (ISINT) d[k] = (T)((UT)a ^ (UT)v); break;              \
            default: return HJO_UNSUPPORTED;                                              \
            }                                                                             \
        }                                                                                 \
        return HJO_OK;                                                                    \
    }
    switch (ty) {
    case HJ_U32: SR_LOOP(uint32_t, uint32_t, 1)
    case HJ_I32: SR_LOOP(int32_t, uint32_t, 1)
    case HJ_U64: SR_LOOP(uint64_t, uint64_t, 1)
    case HJ_I64: SR_LOOP(int64_t, uint64_t, 1)
    case HJ_F32:
        if (op == HJ_REDUCE_OR || op == HJ_REDUCE_AND || op == HJ_REDUCE_XOR) return HJO_UNSUPPORTED;
        SR_LOOP(float, uint32_t, 0)
    default: return HJO_UNSUPPORTED;
    }
#undef SR_LOOP
}

/* Multi-threaded u32 sum histogram (per-thread private bins, merged); CPU baseline only. */
int hjo_histogram_u32_mt(size_t n, const uint32_t* idx, uint32_t* dst, size_t n_dst) {
    int nt = NTHREADS();
    uint32_t* priv = (uint32_t*)calloc((size_t)nt * n_dst, sizeof(uint32_t));
    if (!priv) return HJO_OOM;
    int bad = 0;
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        uint32_t* h = priv + (size_t)t * n_dst;
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            uint32_t k = idx[i];
            if (k >= n_dst) { bad = 1; continue; }
            h[k]++;
        }
    }
    for (int t = 0; t < nt; t++)
        for (size_t k = 0; k < n_dst; k++) dst[k] += priv[(size_t)t * n_dst + k];
    free(priv);
    return bad ? HJO_INVALID : HJO_OK;
}

int hjo_gather(size_t elem_bytes, size_t n, const void* src, size_t n_src, const uint32_t* idx,
               void* dst) {
    if (!src || !idx || !dst) return HJO_INVALID;
    const unsigned char* s = (const unsigned char*)src;
    unsigned char* d = (unsigned char*)dst;
    for (size_t i = 0; i < n; i++) {
        if (idx[i] >= n_src) return HJO_INVALID;
        memcpy(d + i * elem_bytes, s + (size_t)idx[i] * elem_bytes, elem_bytes);
    }
    return HJO_OK;
}

/* =====================================================================================
 * BASELINE.json config C2, the canonical fused elementwise trace:
 *     t = fma(x, 1.5, 0.25);  y = select(x > 0, sin(t), exp2(t))
 * evaluated per element with the per-op meaning of codegen/glsl/mod.rs:688-694 (Select),
 * :830-835 (Gt), :905-910 (Sin, Exp2) and FMA as IEEE fused multiply-add (deviation D1: the
 * reference emits nothing for FMA, glsl/mod.rs:913).  Transcendentals are evaluated in
 * double and rounded once, i.e. the correctly-rounded f32 answer up to double rounding;
 * the GPU is compared against it within the tolerance stated in tests/ (the Vulkan
 * precision table is the reference's own bound: sin abs 2^-11, exp2 3+2|x| ULP).
 * Multi-threaded so that it can serve as the CPU baseline of the bench.
 * ===================================================================================== */
int hjo_c2_chain_f32(size_t n, const float* x, float* y) {
    if (!x || !y) return HJO_INVALID;
#pragma omp parallel for schedule(static) num_threads(NTHREADS())
    for (long long i = 0; i < (long long)n; i++) {
        float xi = x[i];
        float t = fmaf(xi, 1.5f, 0.25f);
        y[i] = xi > 0.f ? (float)sin((double)t) : (float)exp2((double)t);
    }
    return HJO_OK;
}
/* same chain with the host libm's float functions (what a CPU Vulkan ICD would execute);
 * used only for the timed CPU baseline so it is not handicapped by double evaluation. */
int hjo_c2_chain_f32_fast(size_t n, const float* x, float* y) {
    if (!x || !y) return HJO_INVALID;
#pragma omp parallel for schedule(static) num_threads(NTHREADS())
    for (long long i = 0; i < (long long)n; i++) {
        float xi = x[i];
        float t = fmaf(xi, 1.5f, 0.25f);
        y[i] = xi > 0.f ? sinf(t) : exp2f(t);
    }
    return HJO_OK;
}
