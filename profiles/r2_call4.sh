#!/bin/bash
# Round 2, GPU call 4: parity + workloads after the warp-cooperative coarse pass for long reads (K2).
mkdir -p gpurun_out
{
  echo "== parity"
  timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -5
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call4.log 2>&1
tail -30 gpurun_out/r2_call4.log
