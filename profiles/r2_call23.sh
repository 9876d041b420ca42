#!/bin/bash
# call 23: -a N on one input through the single-pass kernel's index + verdict pass (kf_fused<CH,3>) + k2_trim_route<true>
# + K3, against K1 + k2_trim_only + k2_trim_route<true> + K3 (same GPU, alternating); the whole GPU suite.
cd /root/repo
L=gpurun_out/r2_call23.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_nlsave.so build/lib_hybrid.so build/lib_hybrid.so@SICKLE_B200_PATH=general
  echo "== stage times"
  python profiles/workloads.py --general-only
  echo "== pytest -m gpu"
  timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
} > $L 2>&1
tail -40 $L | cut -c1-400
