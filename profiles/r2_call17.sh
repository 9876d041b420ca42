#!/bin/bash
# call 17: K3 with destination-chunk ownership (all loads of a record in flight) vs the run-by-run copy; the whole
# GPU suite (the previous final call stopped at a failing runner case with -x).
cd /root/repo
L=gpurun_out/r2_call17.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B general path, -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_base.so build/lib_k3_new.so build/lib_k3_4.so
  echo "== stage times (shipped lib = k3_new)"
  python profiles/workloads.py --general-only
  echo "== pytest -m gpu"
  timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40
} > $L 2>&1
tail -60 $L | cut -c1-400
