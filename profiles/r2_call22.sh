#!/bin/bash
# call 22: two files -- PASS 2 takes the tile's newline positions from PASS 1 instead of finding them again; before /
# after in alternation on the same GPU (bench.py --config c3, kernel path), two-file parity, then the seeded soak.
cd /root/repo
L=gpurun_out/r2_call22.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  for r in 1 2 3; do
    for lib in build/lib_before_nlsave.so build/lib_nlsave.so; do
      SICKLE_B200_LIB=$PWD/$lib python bench.py --config c3 --steps 20 --warmup 3 --kernel-only 2>/dev/null | tail -1 |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['stage_ms'])"
    done
  done
  echo "== parity: two files, fixtures, command line"
  timeout 1200 python -m pytest tests/test_cuda_parity.py tests/test_reference_fixtures.py tests/test_cli.py -m gpu -q -x -k "two_files or pe or fixture or golden or unequal" 2>&1 | tail -3
  echo "== soak"
  bash profiles/gpu_soak.sh 1 6
} > $L 2>&1
tail -30 $L | cut -c1-300
