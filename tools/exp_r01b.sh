#!/bin/bash
# Round-1 (session 3) measurement pass: GPU tests, compress / histogram timings, ring timeline,
# ncu captures of the compress ring kernel, 2^30 kernel bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python tools/compress_time.py > gpurun_out/compress_time.txt 2>&1; cat gpurun_out/compress_time.txt
python tools/ring_timeline.py 28 > gpurun_out/timeline.txt 2>&1
python tools/hist_check.py > gpurun_out/hist_check.txt 2>&1; cat gpurun_out/hist_check.txt
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:compress_ring -s 2 -c 1 -o gpurun_out/r01b_compress_p01 python tools/prof_driver.py compress:0.01 28 4 > gpurun_out/ncu_c01.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:compress_ring -s 2 -c 1 -o gpurun_out/r01b_compress_p50 python tools/prof_driver.py compress:0.5 28 4 > gpurun_out/ncu_c50.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:hist_ring -s 2 -c 1 -o gpurun_out/r01b_hist python tools/prof_driver.py hist:65536 28 4 > gpurun_out/ncu_hist.log 2>&1
python tools/kernel_bench.py --log2n 30 > gpurun_out/kernel_bench_2p30.txt 2>&1; tail -14 gpurun_out/kernel_bench_2p30.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
