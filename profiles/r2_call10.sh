#!/bin/bash
# Round 2, GPU call 10: batch-summary counters accumulated per CTA instead of one atomic per warp and tile (A/B, same GPU).
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity"
  timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
  echo "== se"
  python profiles/ab_multi.py $V/lib_r1.so $V/lib_prev.so $S
  echo "== pe interleaved"
  python profiles/ab_multi.py --workload pe $V/lib_r1.so $V/lib_prev.so $S
  echo "== pe -M"
  python profiles/ab_multi.py --workload pem $V/lib_prev.so $S
} > gpurun_out/r2_call10.log 2>&1
tail -30 gpurun_out/r2_call10.log | cut -c1-400
