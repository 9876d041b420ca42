#!/bin/bash
# call 18: K3 destination-chunk emit with unconditional boundary loads and branch-free run selection (3 and 4 CTAs/SM),
# K2 as two kernels for short records too (SICKLE_B200_K2_SPLIT=1); same-GPU A/B on the -a 8 general path.
cd /root/repo
L=gpurun_out/r2_call18.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B general path, -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_base.so build/lib_k3_new.so build/lib_k3_4.so build/lib_k3_new.so@SICKLE_B200_K2_SPLIT=1 build/lib_k3_4.so@SICKLE_B200_K2_SPLIT=1
  echo "== stage times (shipped lib = k3_new), then with K2 split"
  python profiles/workloads.py --general-only
  SICKLE_B200_K2_SPLIT=1 python profiles/workloads.py --general-only
  SICKLE_B200_LIB=$PWD/build/lib_k3_4.so python profiles/workloads.py --general-only
  echo "== parity on the general path, both K2 forms"
  SICKLE_B200_K2_SPLIT=1 timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_reference_fixtures.py -m gpu -q -x 2>&1 | tail -3
} > $L 2>&1
tail -40 $L | cut -c1-400
