#!/bin/bash
# call 24: -a N (N <= 32), single end, all on the single-pass kernel (index + verdict pass, kfo_scan / kfo_bases, ordered
# emit pass) against the index pass + k2_trim_route<true> + K3 of call 23 (same GPU, alternating); the single-end kernel
# before / after (its source changed); the whole GPU suite.
cd /root/repo
L=gpurun_out/r2_call24.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_hybrid.so build/lib_ordered.so build/lib_ordered.so@SICKLE_B200_ORDERED=0
  echo "== A/B single end"
  python profiles/ab_multi.py --workload se --rounds 5 --steps 24 build/lib_hybrid.so build/lib_ordered.so
  echo "== pytest -m gpu"
  timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
} > $L 2>&1
tail -40 $L | cut -c1-400
