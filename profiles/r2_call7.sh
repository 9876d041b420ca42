#!/bin/bash
# Round 2, GPU call 7: the new bench.py (all configs at N = 1, reference arm) and configs[3] after the coalesced -n scan.
mkdir -p gpurun_out
{
  echo "== configs[3] K2 by flags"
  python profiles/workloads.py --c4-only
  python profiles/workloads.py --c4-only -x -n
  echo "== sanitize_run (small cases vs the oracle, no tool: compute-sanitizer is closed on this pool)"
  timeout 300 python profiles/sanitize_run.py 2>&1 | tail -22
} > gpurun_out/r2_call7.log 2>&1
for c in c2 c3 c3m c4; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  echo "== bench $c rc=$?" >> gpurun_out/r2_call7.log
  tail -c 3000 gpurun_out/r2_bench_$c.json >> gpurun_out/r2_call7.log
  tail -5 gpurun_out/r2_bench_$c.err >> gpurun_out/r2_call7.log
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "== bench reference rc=$?" >> gpurun_out/r2_call7.log
cat gpurun_out/r2_bench_ref.json >> gpurun_out/r2_call7.log
tail -60 gpurun_out/r2_call7.log | cut -c1-1500
