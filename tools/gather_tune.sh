#!/bin/bash
# Gather with a DRAM-resident table: vectors per thread, L1 bypass, CTAs per SM, L2 fetch granularity.
for fetch in 64 32; do
  for cfg in 20 21 41 81; do
    for ctas in 16 8; do
      echo "== HJ_L2_FETCH=$fetch HJ_GATHER_CFG=$cfg HJ_GATHER_CTAS=$ctas"
      HJ_L2_FETCH=$fetch HJ_GATHER_CFG=$cfg HJ_GATHER_CTAS=$ctas timeout 120 python tools/gather_time.py 2>&1 | grep "gather table"
    done
  done
done
