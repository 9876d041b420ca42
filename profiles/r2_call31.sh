#!/bin/bash
# call 31: the command line file to file in the other modes (24 M reads on /dev/shm): -a 16 (reference order: ordered emit),
# two input files, two input files with -a 16.
cd /root/repo
L=gpurun_out/r2_call31.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  python profiles/cli_bench.py --reads 24000000 --skip-ref --repeat 2 --threads 16
  python profiles/cli_bench.py --reads 8000000 --skip-ref --repeat 2 --two-files
  python profiles/cli_bench.py --reads 8000000 --skip-ref --repeat 2 --two-files --threads 16
} > $L 2>&1
cut -c1-700 $L | tail -8
