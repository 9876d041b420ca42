#!/bin/bash
# Round 2, GPU call 14: 11 KB tiles (CH = 3) for short records; bench.py with its default arguments after the plan fix.
mkdir -p gpurun_out
{
  echo "== parity"
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
  echo "== workloads"
  python profiles/workloads.py
  echo "== se, previous library vs this one"
  python profiles/ab_multi.py build/variants/lib_prev2.so sickle_b200/libsickle_b200.so
} > gpurun_out/r2_call14.log 2>&1
python bench.py > gpurun_out/r2_call14_bench_default.json 2> gpurun_out/r2_call14_bench_default.err
echo "== bench default rc=$?" >> gpurun_out/r2_call14.log
tail -c 1500 gpurun_out/r2_call14_bench_default.json >> gpurun_out/r2_call14.log; tail -5 gpurun_out/r2_call14_bench_default.err >> gpurun_out/r2_call14.log
tail -30 gpurun_out/r2_call14.log | cut -c1-330
