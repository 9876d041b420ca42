#!/bin/bash
# Round 2, GPU call 2: parity of the coarse-to-fine trimmer on the device, A/B against the round-1 library,
# knock-outs, phase timing, one ncu full capture of the new kf_fused<9>.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity"
  timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -5
  echo "== A/B vs round 1, knock-outs"
  python profiles/ab_multi.py $V/lib_r1.so $S $V/lib_KO_S6.so $V/lib_KO_S8A.so $V/lib_KO_FLUSH.so $V/lib_KO_LB1.so $V/lib_KO_LB2.so
  echo "== CH=7"
  SICKLE_B200_FUSED_CH=7 python profiles/ab_multi.py $V/lib_r1.so $S
  echo "== pe interleaved"
  python profiles/ab_multi.py --workload pe $V/lib_r1.so $S
  echo "== phase timing"
  cp $V/lib_timing.so sickle_b200/libsickle_b200_timing.so
  python profiles/phase_timing.py
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call2.log 2>&1
python bench.py --steps 30 --warmup 3 --kernel-only > gpurun_out/r2_call2_bench.json 2> gpurun_out/r2_call2_bench.err
ncu --set full --clock-control none --import-source on -k regex:kf_fused -s 4 -c 1 -o gpurun_out/r2_v8 -f python bench.py --steps 6 --warmup 3 --kernel-only > gpurun_out/r2_call2_ncu.log 2>&1
ncu -i gpurun_out/r2_v8.ncu-rep --page raw --csv > gpurun_out/r2_v8_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_v8.ncu-rep --page source --csv > gpurun_out/r2_v8_source.csv 2>/dev/null
tail -60 gpurun_out/r2_call2.log
