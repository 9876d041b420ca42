#!/bin/bash
# call 28: long-record batches with K3 as a kernel of its own (48 registers) against the branch inside k3_emit (80).
cd /root/repo
L=gpurun_out/r2_call28.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  for r in 1 2 3; do
    for lib in build/lib_final2.so build/lib_k3long.so; do
      echo "-- $lib"
      SICKLE_B200_LIB=$PWD/$lib python profiles/workloads.py --c4-only -x -n | cut -c1-330
    done
  done
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "long or golden or sizes" 2>&1 | tail -3
} > $L 2>&1
tail -20 $L | cut -c1-330
