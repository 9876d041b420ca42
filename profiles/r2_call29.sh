#!/bin/bash
# call 29: k2_trim_only at 5 and 6 CTAs per SM (48 / 40 registers, some spills) against 4 (58 registers) on configs[3].
cd /root/repo
L=gpurun_out/r2_call29.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  for r in 1 2 3; do
    for lib in build/lib_k3long.so build/lib_k2a5.so build/lib_k2a6.so; do
      echo "-- $lib"
      SICKLE_B200_LIB=$PWD/$lib python profiles/workloads.py --c4-only -x -n | cut -c1-330
    done
  done
} > $L 2>&1
grep -E "^--|stage_ms" $L | sed 's/"workload".*"stage_ms"/stage_ms/' | cut -c1-120
