#!/bin/bash
# call 25: five CTAs per SM (48 registers, no spill) instead of four for the passes that stage nothing (PASS 1 and 3):
# -a 8 in one process; two files by alternating bench runs.
cd /root/repo
L=gpurun_out/r2_call25.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  echo "== A/B -a 8, 1 M reads"
  python profiles/ab_multi.py --workload a8 --rounds 5 --steps 12 build/lib_ordered.so build/lib_p1x5.so
  echo "== two files"
  for r in 1 2 3; do
    for lib in build/lib_ordered.so build/lib_p1x5.so; do
      SICKLE_B200_LIB=$PWD/$lib python bench.py --config c3 --steps 20 --warmup 3 --kernel-only 2>/dev/null | tail -1 |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['stage_ms'])"
    done
  done
  echo "== parity with the 5-CTA build"
  SICKLE_B200_LIB=$PWD/build/lib_p1x5.so timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "two_files or pe or thread_order or golden" 2>&1 | tail -3
} > $L 2>&1
tail -30 $L | cut -c1-300
