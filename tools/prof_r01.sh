#!/bin/bash
# Round-1 profiling pass (run under gpurun): launch list of the default bench command plus one
# `ncu --set full` capture per hot kernel.  Outputs land in gpurun_out/ (summaries are then made
# here with tools/ncu_summary.py and committed under profiles/).
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_under_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:hj_kernel -s 3 -c 2 -o gpurun_out/r01_fused python bench.py --steps 3 --warmup 3 --no-suite > gpurun_out/ncu_fused.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:scan_ring -s 2 -c 1 -o gpurun_out/r01_scan python tools/prof_driver.py scan 28 4 > gpurun_out/ncu_scan.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:compress_ring -s 2 -c 1 -o gpurun_out/r01_compress_p50 python tools/prof_driver.py compress:0.5 28 4 > gpurun_out/ncu_compress.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:compress_ring -s 2 -c 1 -o gpurun_out/r01_compress_p01 python tools/prof_driver.py compress:0.01 28 4 > gpurun_out/ncu_compress01.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:reduce_kernel -s 2 -c 1 -o gpurun_out/r01_reduce python tools/prof_driver.py reduce 28 4 > gpurun_out/ncu_reduce.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:hist_ring -s 2 -c 1 -o gpurun_out/r01_hist python tools/prof_driver.py hist:65536 28 4 > gpurun_out/ncu_hist.log 2>&1
ls -la gpurun_out
