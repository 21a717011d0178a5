/* abi_smoke.c — a pure C11 consumer of include/hj.h: no Python, no C++.
 *
 * Proves that the header is valid C, that libhj_b200.so can be driven by plain FFI calls, and walks
 * the call sequence the Rust binding makes (bindings/rust/backend_cuda.rs; the reference side is
 * hephaestus-jit/src/backend/mod.rs:26-49 and graph.rs:315-323):
 *   Device::cuda(0) -> create_buffer_from_slice -> execute_graph(passes, env) -> to_host
 * with a hand-built IR (the IR `Compiler::compile` emits for y = x * 3 + 1 on u32, compiler.rs:20-65),
 * a Reduce, a PrefixSum and a Compress pass, all checked against values computed here.
 *
 * Exit code 0: everything matched.  77: no CUDA device (the build check still proved the header).
 * Built by __graft_entry__.build(); run by tests/test_abi_c_gpu.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hj.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        hj_status s_ = (call);                                                        \
        if (s_ != HJ_OK) {                                                            \
            fprintf(stderr, "%s failed (%d): %s\n", #call, (int)s_, hj_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

enum { N = (1 << 20) + 5 };

int main(void) {
    printf("hj abi version %u, %d device(s)\n", hj_abi_version(), (int)hj_device_count());
    if (hj_device_count() == 0) {
        hj_device* none = NULL;
        hj_status s = hj_device_create(0, &none);  /* must fail loudly: there is no CPU fallback */
        if (s != HJ_ERR_NO_DEVICE) {
            fprintf(stderr, "expected HJ_ERR_NO_DEVICE without a GPU, got %d\n", (int)s);
            return 1;
        }
        printf("no CUDA device: %s\n", hj_last_error());
        return 77;
    }
    hj_device* dev = NULL;
    CHECK(hj_device_create(0, &dev));

    uint32_t* x = malloc(sizeof(uint32_t) * N);
    uint8_t* mask = malloc(N);
    uint32_t* got = malloc(sizeof(uint32_t) * N);
    uint32_t seed = 12345u;
    for (size_t i = 0; i < N; i++) {
        seed = seed * 1664525u + 1013904223u;
        x[i] = seed >> 12;
        mask[i] = (uint8_t)((seed >> 7) & 1u);
    }

    /* resources: 0 x, 1 y, 2 sum, 3 scan, 4 mask, 5 index, 6 count */
    hj_buffer* env[7] = {0};
    CHECK(hj_buffer_create_from_slice(dev, x, sizeof(uint32_t) * N, &env[0]));
    CHECK(hj_buffer_create(dev, sizeof(uint32_t) * N, &env[1]));
    CHECK(hj_buffer_create(dev, sizeof(uint32_t), &env[2]));
    CHECK(hj_buffer_create(dev, sizeof(uint32_t) * N, &env[3]));
    CHECK(hj_buffer_create_from_slice(dev, mask, N, &env[4]));
    CHECK(hj_buffer_create(dev, sizeof(uint32_t) * N, &env[5]));
    CHECK(hj_buffer_create(dev, sizeof(uint32_t), &env[6]));
    CHECK(hj_buffer_fill_zero(env[5]));
    const hj_buffer_desc descs[7] = {{N, HJ_U32, 4}, {N, HJ_U32, 4}, {1, HJ_U32, 4}, {N, HJ_U32, 4},
                                     {N, HJ_BOOL, 1}, {N, HJ_U32, 4}, {1, HJ_U32, 4}};

    /* IR of  y[i] = x[i] * 3 + 1  (types: 0 = U32, 1 = Void) */
    const hj_type_desc types[2] = {{HJ_U32, 0, 0, 0, 0, 0}, {HJ_VOID, 0, 0, 0, 0, 0}};
    const uint32_t deps[] = {0, 1, /* gather */ 2, 3, /* mul */ 4, 5, /* add */ 7, 6, 1 /* scatter */};
    const hj_ir_var vars[] = {
        {0, HJ_OP_BUFFER_REF, 0, 0, 0, 0, 0},          /* 0: x                */
        {0, HJ_OP_INDEX, 0, 0, 0, 0, 0},               /* 1: index            */
        {0, HJ_OP_GATHER, 0, 0, 2, 0, 0},              /* 2: x[index]         */
        {0, HJ_OP_LITERAL, 0, 0, 0, 0, 3},             /* 3: 3u               */
        {0, HJ_OP_BOP, HJ_BOP_MUL, 2, 4, 0, 0},        /* 4: x * 3            */
        {0, HJ_OP_LITERAL, 0, 0, 0, 0, 1},             /* 5: 1u               */
        {0, HJ_OP_BOP, HJ_BOP_ADD, 4, 6, 0, 0},        /* 6: x * 3 + 1        */
        {0, HJ_OP_BUFFER_REF, 0, 0, 0, 0, 1},          /* 7: y                */
        {1, HJ_OP_SCATTER, 0, 6, 9, 0, 0},             /* 8: y[index] = var 6 */
    };
    const hj_ir ir = {vars, 9, deps, 9, types, 2, NULL, 0, 2};
    char* source = NULL;
    CHECK(hj_ir_codegen(&ir, &source));
    printf("generated CUDA C++: %zu bytes, ir hash %016llx\n", strlen(source), (unsigned long long)hj_ir_hash(&ir));
    hj_free_string(source);

    const uint32_t r_kernel[] = {0, 1}, r_reduce[] = {2, 1}, r_scan[] = {3, 0}, r_compress[] = {5, 6, 4};
    const hj_pass passes[4] = {
        {HJ_PASS_KERNEL, 0, r_kernel, 2, -1, &ir, N},
        {HJ_PASS_REDUCE, HJ_REDUCE_SUM, r_reduce, 2, -1, NULL, 0},
        {HJ_PASS_PREFIX_SUM, 1, r_scan, 2, -1, NULL, 0},
        {HJ_PASS_COMPRESS, 0, r_compress, 3, -1, NULL, 0},
    };
    hj_pass_report pr[4];
    hj_report report = {0.0, 0, pr, 4};
    CHECK(hj_execute_graph(dev, passes, 4, env, descs, 7, &report));
    for (uint32_t i = 0; i < report.n_passes; i++) printf("  pass %u %-24s %8.1f us\n", i, pr[i].name, pr[i].duration_us);
    /* and once more through the relaunch path (second launch captures, third replays) */
    uint32_t how = 9;
    for (int rep = 0; rep < 3; rep++) CHECK(hj_execute_graph_cached(dev, 42, passes, 4, env, descs, 7, &how));
    printf("relaunch path: how = %u (2 = replayed one captured CUDA graph)\n", how);

    int bad = 0;
    uint32_t sum = 0, run = 0, count = 0, scalar = 0;
    CHECK(hj_buffer_to_host(env[1], 0, sizeof(uint32_t) * N, got));
    for (size_t i = 0; i < N; i++) {
        const uint32_t want = x[i] * 3u + 1u;
        if (got[i] != want && bad++ < 5) fprintf(stderr, "y[%zu] = %u, expected %u\n", i, got[i], want);
        sum += want;
    }
    CHECK(hj_buffer_to_host(env[2], 0, sizeof(uint32_t), &scalar));
    if (scalar != sum) { fprintf(stderr, "sum = %u, expected %u\n", scalar, sum); bad++; }
    CHECK(hj_buffer_to_host(env[3], 0, sizeof(uint32_t) * N, got));
    for (size_t i = 0; i < N; i++) {
        run += x[i];
        if (got[i] != run && bad++ < 5) fprintf(stderr, "scan[%zu] = %u, expected %u\n", i, got[i], run);
    }
    CHECK(hj_buffer_to_host(env[5], 0, sizeof(uint32_t) * N, got));
    for (size_t i = 0; i < N; i++)
        if (mask[i]) {
            if (got[count] != (uint32_t)i && bad++ < 5) fprintf(stderr, "index[%u] = %u, expected %zu\n", count, got[count], i);
            count++;
        }
    for (size_t i = count; i < N; i++)
        if (got[i] != 0 && bad++ < 5) fprintf(stderr, "index[%zu] beyond the count was touched\n", i);
    CHECK(hj_buffer_to_host(env[6], 0, sizeof(uint32_t), &scalar));
    if (scalar != count) { fprintf(stderr, "count = %u, expected %u\n", scalar, count); bad++; }

    /* error behaviour: a status and a message, never a crash */
    if (hj_reduce(dev, HJ_REDUCE_SUM, HJ_U32, (size_t)N * 2, env[0], env[2]) != HJ_ERR_INVALID) { fprintf(stderr, "oversized reduce was accepted\n"); bad++; }
    if (hj_reduce(dev, HJ_REDUCE_AND, HJ_F32, N, env[0], env[2]) != HJ_ERR_UNSUPPORTED) { fprintf(stderr, "reduce(And, F32) must be unsupported (todo!() in reduce.rs)\n"); bad++; }

    uint64_t launches = 0;
    CHECK(hj_device_launch_count(dev, &launches));
    for (int i = 0; i < 7; i++) CHECK(hj_buffer_release(env[i]));
    CHECK(hj_device_release(dev));
    free(x); free(mask); free(got);
    printf("%s: %llu kernel launches, %d mismatches\n", bad ? "FAILED" : "ok", (unsigned long long)launches, bad);
    return bad ? 1 : 0;
}
