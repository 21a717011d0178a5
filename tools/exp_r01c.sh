#!/bin/bash
# Round-1 (session 4) measurement pass: full GPU test suite, bench line (own + reference arm), ncu launch
# list of the default bench command, ncu capture of the histogram kernel, 2^30 kernel bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01c_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r01c_pytest_gpu.txt
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r01c_bench_reference_arm.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/r01c_bench_reference_arm.json
python bench.py > gpurun_out/r01c_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1200 gpurun_out/r01c_bench_n1.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01c_launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r01c_bench_under_ncu.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:hist_ring -s 2 -c 1 -o gpurun_out/r01c_hist python tools/prof_driver.py hist:65536 28 4 > gpurun_out/ncu_hist.log 2>&1
python tools/kernel_bench.py --log2n 30 > gpurun_out/r01c_kernel_bench_2p30.txt 2>&1; tail -16 gpurun_out/r01c_kernel_bench_2p30.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
