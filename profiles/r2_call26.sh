#!/bin/bash
# call 26: -a 8 with the offsets in one kernel (group sums by atomics in the index pass) and five CTAs per SM for the
# passes that stage nothing, against call 24's library; parity.
cd /root/repo
L=gpurun_out/r2_call26.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_ordered.so build/lib_final2.so
  echo "== stages"
  python profiles/workloads.py --a8-only
  echo "== parity"
  timeout 1500 python -m pytest tests/test_cuda_parity.py tests/test_reference_fixtures.py tests/test_cli.py -m gpu -q --tb=short 2>&1 | tail -8
} > $L 2>&1
tail -30 $L | cut -c1-300
