#!/bin/bash
# call 27: -a 32 (ordered emit) and -a 33 (index pass + K2 routing + K3) against the oracle; seeded soak, seeds 10-17.
cd /root/repo
L=gpurun_out/r2_call27.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q --tb=short -k "thread_order" 2>&1 | tail -5
  bash profiles/gpu_soak.sh 10 8
} > $L 2>&1
tail -30 $L | cut -c1-300
