#!/bin/bash
# Round 2, GPU call 11: ticket poisoning instead of a per-tile read of the Control block; K2's counters per CTA. A/B + ncu.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity"
  timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
  echo "== se"
  python profiles/ab_multi.py $V/lib_r1.so $V/lib_prev.so $S
  echo "== se CH=7"
  SICKLE_B200_FUSED_CH=7 python profiles/ab_multi.py $V/lib_r1.so $S
  echo "== pe interleaved"
  python profiles/ab_multi.py --workload pe $V/lib_r1.so $S
  echo "== -a 8 (general path)"
  python profiles/ab_multi.py --workload a8 $V/lib_r1.so $V/lib_prev.so $S
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call11.log 2>&1
python bench.py --steps 20 --warmup 3 --kernel-only > gpurun_out/r2_call11_bench.json 2> gpurun_out/r2_call11_bench.err
ncu --set full --clock-control none --import-source on -k regex:kf_fused -s 4 -c 1 -o gpurun_out/r2_v9 -f python bench.py --steps 6 --warmup 3 --kernel-only --min-timed-s 0.01 > gpurun_out/r2_call11_ncu.log 2>&1
ncu -i gpurun_out/r2_v9.ncu-rep --page raw --csv > gpurun_out/r2_v9_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_v9.ncu-rep --page source --csv > gpurun_out/r2_v9_source.csv 2>/dev/null
tail -40 gpurun_out/r2_call11.log | cut -c1-330
