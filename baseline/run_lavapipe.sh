#!/bin/bash
# The reference's own path on a CPU Vulkan ICD (BASELINE.json config 1) — for an image that HAS the
# toolchain.  This repository's build image has none of it (no cargo / nightly rustc, no Vulkan loader,
# no Mesa lavapipe ICD, no shaderc / glslang, no network: SURVEY.md §8c), so this script is not run here;
# bench.py's reference arm times the oracle's C port of the same algorithms on the host cores instead
# and says so in its JSON line ("kind": "port").
#
# Needs: rustup toolchain nightly; libvulkan1 + mesa-vulkan-drivers (lvp_icd.x86_64.json);
# vulkan-validationlayers (the reference enables the Khronos validation layer unconditionally,
# vulkan_core/device.rs:104-105); cmake + python3 for the shaderc / glslang builds.
#
# Bench ids: hephaestus-jit/benches/vulkan.rs:157-192 — groups "prefix_sum_large_u32" and
# "compress_large", parameter = n (2^10 .. 2^29).  The reference has no reduce bench; config 1's f32
# sum needs a `reduce_sum` group added next to them (x = sized_literal(1f32, n); x.reduce_sum()).
# Note defect D5: the reduce shader has no barrier in its tree loop and is only correct
# when a 32-invocation workgroup runs in lock step, which lavapipe does not guarantee.
set -euo pipefail
REF=${1:-/root/reference}
export VK_ICD_FILENAMES=${VK_ICD_FILENAMES:-/usr/share/vulkan/icd.d/lvp_icd.x86_64.json}
export LP_NUM_THREADS=${LP_NUM_THREADS:-$(nproc)}
echo "lavapipe on ${LP_NUM_THREADS} host threads" >&2
cd "$REF/hephaestus-jit"
cargo +nightly bench --bench vulkan -- 'prefix_sum_large_u32/1048576|compress_large/1048576'
