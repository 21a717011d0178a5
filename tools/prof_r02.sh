#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU): the launch list of the default bench command and
# one `ncu --set full` capture per hot kernel at the size bench.py runs it.  The .ncu-rep files land in
# gpurun_out/; tools/prof_r02_summarise.sh turns them into profiles/r02_ncu_summary.txt,
# profiles/r02_launches_bench_summary.txt and profiles/traffic.json (read by bench.py).
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
cap() {  # name regex driver-args...
  local name=$1 regex=$2; shift 2
  timeout 300 $NCU --set full --import-source on -k regex:$regex -s 2 -c 1 -o gpurun_out/r02_$name python tools/prof_driver.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
cap map hj_kernel_vec map 28 4
cap reduce reduce_kernel reduce 30 4
cap scan scan_ring scan 30 4
cap compress_p50 compress_ring compress:0.5 30 4
cap compress_p01 compress_ring compress:0.01 30 4
cap compress_p99 compress_ring compress:0.99 30 4
cap hist hist_ring hist:65536 28 4
cap hist_fold hist_fold hist:65536 28 4
cap gather_dram gather4 gather:28 28 4
cap gather_l2 gather4 gather:20 28 4
# the two DynSize kernels of the wavefront step (profiles/r02_wavefront.md): launches 2 and 3 of the NVRTC entry
timeout 300 $NCU --set full --import-source on -k regex:hj_kernel -s 2 -c 2 -o gpurun_out/r02_wavefront python tools/prof_driver.py wavefront 28 3 > gpurun_out/ncu_wavefront.log 2>&1
ls -la gpurun_out
