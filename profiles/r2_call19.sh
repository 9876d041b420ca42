#!/bin/bash
# call 19: long-read K2 with the round-level fast paths (two reductions per 128 steps; one vote per 512 bases of the
# -n scan) against the library before them, alternating on the same GPU; ncu of the new K3.
cd /root/repo
L=gpurun_out/r2_call19.log
{
  nvidia-smi --query-gpu=name,clocks.sm,power.limit --format=csv,noheader
  for r in 1 2 3; do
    for lib in build/lib_k3_new.so build/lib_k2fast.so; do
      echo "-- $lib"
      SICKLE_B200_LIB=$PWD/$lib python profiles/workloads.py --c4-only -x -n | cut -c1-330
    done
  done
  echo "== parity, long reads and the rest of the general path"
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "long or fuzz or golden or sizes or kernel" 2>&1 | tail -3
} > $L 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k3_emit" -s 2 -c 1 -o gpurun_out/r2_k3v2 -f python profiles/workloads.py --general-only > gpurun_out/r2_call19_ncu.log 2>&1
ncu -i gpurun_out/r2_k3v2.ncu-rep --page raw --csv > gpurun_out/r2_k3v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_k3v2.ncu-rep --page source --csv > gpurun_out/r2_k3v2_source.csv 2>/dev/null
tail -40 $L | cut -c1-400
