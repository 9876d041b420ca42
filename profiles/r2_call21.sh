#!/bin/bash
# call 21: long-read K2 with 256-step rounds and one unit per ticket vs the library before (alternating, same GPU);
# then the command line on 100 M reads (configs[1] at full size), file to file on /dev/shm.
cd /root/repo
L=gpurun_out/r2_call21.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  for r in 1 2 3; do
    for lib in build/lib_k2fast.so build/lib_k2u8.so; do
      echo "-- $lib"
      SICKLE_B200_LIB=$PWD/$lib python profiles/workloads.py --c4-only -x -n | cut -c1-330
    done
  done
  echo "== parity, long reads"
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "long or fuzz or golden" 2>&1 | tail -3
  echo "== command line, 100 M reads"
  df -h /dev/shm | tail -1
  timeout 900 python profiles/cli_bench.py --reads 100000000 --skip-ref --repeat 2
} > $L 2>&1
tail -30 $L | cut -c1-600
