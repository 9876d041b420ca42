#!/bin/bash
# Round 2, first GPU call: knock-outs and the queued build variants, timed in one process each (profiles/ab_multi.py),
# then the other workload shapes incl. configs[3].  Libraries are prebuilt under build/variants/.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== knock-outs (CH adaptive -> 9)"
  python profiles/ab_multi.py $S $V/lib_KO_S6.so $V/lib_KO_S8A.so $V/lib_KO_FLUSH.so $V/lib_KO_LB1.so $V/lib_KO_LB2.so $V/lib_KO_ALL.so
  echo "== variants"
  python profiles/ab_multi.py $S $V/lib_SPLIT4.so $V/lib_LONGFIRST.so $V/lib_NLBF.so $V/lib_COMBO3.so $V/lib_DIRECT.so $V/lib_DIRECT_EARLY.so
  echo "== CH=7 pinned"
  SICKLE_B200_FUSED_CH=7 python profiles/ab_multi.py $S $V/lib_DIRECT.so $V/lib_DIRECT_EARLY.so $V/lib_COMBO3.so
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call1.log 2>&1
tail -70 gpurun_out/r2_call1.log
