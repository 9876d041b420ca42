#!/bin/bash
# Round 2, GPU call 3: (a) K2/K3 for long records (configs[3]); (b) the emit-now experiment (no staging buffer,
# immediate look-back #2, records copied straight to global memory) at 3 / 4 / 5 CTAs per SM.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity"
  timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -5
  echo "== emit-now, adaptive tile (CH 9)"
  python profiles/ab_multi.py $S $V/lib_EN3.so $V/lib_EN4.so $V/lib_EN5.so
  echo "== CH=11"
  SICKLE_B200_FUSED_CH=11 python profiles/ab_multi.py $S $V/lib_EN4.so
  echo "== CH=7"
  SICKLE_B200_FUSED_CH=7 python profiles/ab_multi.py $S $V/lib_EN4.so $V/lib_EN5.so
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call3.log 2>&1
tail -40 gpurun_out/r2_call3.log
