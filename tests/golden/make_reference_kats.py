#!/usr/bin/env python3
"""Transcribes the known-answer vectors the reference's own tests hold for the hot path
(hephaestus-jit/src/test.rs) into tests/golden/reference_kats.json.

The reference cannot run in this image (nightly Rust + Vulkan), so these are transcriptions
of the inputs/expected outputs written in its test source, each with the file:line it comes
from.  Deterministic: re-running this script reproduces the committed JSON byte for byte.
Cases whose reference input is `rand::thread_rng()` (seedless) are NOT here; the tests
re-create them with fixed numpy seeds and the host fold the reference compares against.
"""
import json
import os

kats = {
    "_about": "known-answer tests transcribed from hephaestus-jit/src/test.rs (see 'src' of each case)",
    "reduce": [
        # reduce_max, test.rs:493-518
        {"src": "test.rs:515", "op": "max", "ty": "U8", "range": [0, 255], "expect": 254},
        {"src": "test.rs:516", "op": "max", "ty": "I8", "range": [-128, 127], "expect": 126},
        {"src": "test.rs:517", "op": "max", "ty": "F32", "range": [0, 100], "expect": 99.0},
        # reduce_min, test.rs:519-546
        {"src": "test.rs:541", "op": "min", "ty": "U8", "range": [0, 255], "expect": 0},
        {"src": "test.rs:542", "op": "min", "ty": "I8", "range": [-128, 127], "expect": -128},
        {"src": "test.rs:543", "op": "min", "ty": "I64", "range": [-128, 127], "expect": -128},
        {"src": "test.rs:544", "op": "min", "ty": "U64", "range": [0, 65535], "expect": 0},
        {"src": "test.rs:545", "op": "min", "ty": "F32", "range": [0, 100], "expect": 0.0},
    ],
    # reduce_sum / reduce_prod / reduce_and / reduce_or / reduce_xor (test.rs:548-802):
    # 1000 random values, compared with the wrapping host fold.  (ty, lo, hi) as in the source.
    "reduce_random": [
        {"src": "test.rs:576", "op": "sum", "ty": "U8", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:577", "op": "sum", "ty": "I8", "lo": -128, "hi": 127, "n": 1000},
        {"src": "test.rs:578", "op": "sum", "ty": "I64", "lo": -128, "hi": 127, "n": 1000},
        {"src": "test.rs:579", "op": "sum", "ty": "U64", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:580", "op": "sum", "ty": "F32", "lo": 0, "hi": 100, "n": 1000,
         "note": "integer-valued f32: every partial sum is exact, so assert_eq holds"},
        {"src": "test.rs:638", "op": "prod", "ty": "U8", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:639", "op": "prod", "ty": "I8", "lo": -128, "hi": 127, "n": 1000},
        {"src": "test.rs:640", "op": "prod", "ty": "I64", "lo": -128, "hi": 127, "n": 1000},
        {"src": "test.rs:641", "op": "prod", "ty": "U64", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:682-684", "op": "and", "ty": "U8", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:685", "op": "and", "ty": "U64", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:686-700", "op": "and", "ty": "Bool", "lo": 0, "hi": 2, "n": 1000},
        {"src": "test.rs:732", "op": "or", "ty": "U8", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:733", "op": "or", "ty": "U64", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:734-750", "op": "or", "ty": "Bool", "lo": 0, "hi": 2, "n": 1000},
        {"src": "test.rs:782", "op": "xor", "ty": "U8", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:783", "op": "xor", "ty": "U64", "lo": 0, "hi": 255, "n": 1000},
        {"src": "test.rs:784-800", "op": "xor", "ty": "Bool", "lo": 0, "hi": 2, "n": 1000},
    ],
    # reduce_prod f32 (test.rs:642-649): 10 values log2(0.01*k), k in 1..1000, abs eps 0.01
    "reduce_prod_f32": {"src": "test.rs:642-649", "n": 10, "k_lo": 1, "k_hi": 1000, "abs_eps": 0.01},
    # prefix_sum (test.rs:949-975): inclusive u64 scan of 0..8195 (2048*4+3)
    "prefix_sum": {"src": "test.rs:949-975", "ty": "U64", "n": 2048 * 4 + 3, "inclusive": True,
                   "input": "arange", "expect_last": (8195 * 8194) // 2},
    # bench invariant (benches/vulkan.rs:126-137,186-192): prefix_sum(false) of n ones must end
    # in n, i.e. the reference's "exclusive" call is inclusive
    "prefix_sum_bench_invariant": {"src": "benches/vulkan.rs:130-137", "ty": "U32",
                                   "n": [1024, 2048, 4096, 1 << 16], "ref_compat": True},
    # compress_large (test.rs:917-948): n = 2^12 + 15 all true -> indices 0..n, count n
    "compress_all_true": {"src": "test.rs:917-948", "n": 4096 + 15},
    # compress_small (test.rs:885-916): 128 random bools vs host filter
    "compress_small": {"src": "test.rs:885-916", "n": 128},
    # bench invariant (benches/vulkan.rs:106-114): compress of n trues -> count n
    "compress_bench_invariant": {"src": "benches/vulkan.rs:106-114", "n": [1024, 4096, 1 << 16]},
    # scatter_reduce (test.rs:863-883): 16 x (+1u32) into bin 0 of [0,0,0] -> 16
    "scatter_reduce": {"src": "test.rs:863-883", "dst": [0, 0, 0], "n": 16, "idx": 0,
                       "value": 1, "op": "sum", "ty": "U32", "expect": [16, 0, 0]},
    # dynamic_index (test.rs:976-1019): filter 3 < v < 7 over 1024 random ints in 0..10
    "dynamic_index": {"src": "test.rs:976-1019", "n": 1024, "min": 3, "max": 7, "lo": 0, "hi": 10},
    # uop_cos (test.rs:804-828): cos of [0, 1, pi], abs eps 1e-3
    "uop_cos": {"src": "test.rs:804-828", "x": [0.0, 1.0, 3.141592653589793], "abs_eps": 1e-3},
    # elementwise / scatter / gather programs with literal expected outputs
    "programs": {
        "simple1": {"src": "test.rs:62-92", "i": [1, 2, 3, 4, 5, 5, 6, 7, 8, 9], "j": [0, 1, 2, 3, 4]},
        "simple_u16": {"src": "test.rs:94-115", "c": [1] * 10},
        "simple_f16": {"src": "test.rs:116-128", "c": list(range(10))},
        "scatter_chain1": {"src": "test.rs:130-154", "b1": [2, 2, 2, 2, 2]},
        "scatter_chain2": {"src": "test.rs:155-175", "b": [2, 2, 2, 2, 2], "a": [1, 1, 1, 1, 1]},
        "conditional_scatter": {"src": "test.rs:348-372",
                                "active": [1, 1, 0, 0, 1, 0, 1, 0, 1, 0],
                                "dst": [1, 1, 0, 0, 1, 0, 1, 0, 1, 0]},
        "conditional_gather": {"src": "test.rs:373-392",
                               "active": [1, 1, 0, 0, 1, 0, 1, 0, 1, 0],
                               "dst": [1, 1, 0, 0, 1, 0, 1, 0, 1, 0]},
        "select": {"src": "test.rs:393-409", "cond": [1, 0], "res": [10, 5]},
        "array": {"src": "test.rs:1288-1301", "array": [[1, 2, 3], [1, 2, 3]]},
        "vec3_memory_layout": {"src": "test.rs:1322-1336", "vec": [1, 2, 3, 1, 2, 3]},
        "cast_array_vec": {"src": "test.rs:1337-1356", "arr": [1, 2, 3, 1, 2, 3]},
        "if_record1": {"src": "test.rs:1411-1434", "i": [1, 0]},
        "loop_record1": {"src": "test.rs:1436-1462", "i": [2, 2]},
        "loop_record2": {"src": "test.rs:1464-1487", "i": [2, 2], "c": [0, 0]},
        "loop_record_side_effect": {"src": "test.rs:1488-1510",
                                    "dst": [1, 1, 1, 1, 0, 0, 0, 0, 0, 0]},
        "record_test": {"src": "test.rs:1063-1083", "a": [2, 3, 4], "b": [5, 6, 7]},
        "record_output": {"src": "test.rs:1084-1101", "a1": [2, 3, 4]},
        "record_change": {"src": "test.rs:1125-1140", "a1": [2, 3, 4], "a2": [2, 3, 4, 5]},
        "record_fn": {"src": "test.rs:1160-1175", "y": [1, 2, 3, 4]},
        "recorded_change": {"src": "test.rs:1671-1689", "y1": [2, 3, 4], "y2": [2, 3, 4]},
        "aliasing1": {"src": "test.rs:1638-1670", "aliasing_rate": 0.66666666, "eps": 1e-4},
    },
    # struct layout KATs (vartype.rs:387-461): offsets/size must equal #[repr(C)]
    "layout": [
        {"src": "vartype.rs:391-407", "tys": ["U8", "U32"], "offsets": [0, 4], "size": 8},
        {"src": "vartype.rs:408-424", "tys": ["U32", "U8"], "offsets": [0, 4], "size": 8},
        {"src": "vartype.rs:425-451", "tys": ["U8", "U16", "U32", "U64"], "offsets": [0, 2, 4, 8],
         "size": 16},
        {"src": "vartype.rs:452-460", "tys": [["Array", "F32", 12], "U32"], "offsets": [0, 48],
         "size": 52},
    ],
}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
with open(out, "w") as f:
    json.dump(kats, f, indent=1, sort_keys=True)
    f.write("\n")
print("wrote", out)
