#!/bin/bash
# Round 2, GPU call 12: two input files through both passes of the single-pass kernel (configs[2]): parity, then speed
# against the previous library (K1/K2/K3 for two files), and a check that the single-pass kernel itself is unchanged.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity (C ABI, both paths) + reference fixtures + CLI"
  timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_reference_fixtures.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -4
  echo "== se / pe interleaved: previous library vs this one"
  python profiles/ab_multi.py $V/lib_prev2.so $S
  python profiles/ab_multi.py --workload pe $V/lib_prev2.so $S
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call12.log 2>&1
for c in c3 c3m c2; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err
  echo "== bench $c rc=$?" >> gpurun_out/r2_call12.log
  python - <<PY >> gpurun_out/r2_call12.log
import json
d=json.loads(open('gpurun_out/r2b_bench_$c.json').read().strip().splitlines()[-1])
print(d['config']['config_id'], 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), d['roofline']['stage_ms'], 'launches', d['gpu_launches'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,1), d['e2e'] and round(d['e2e']['frac_of_bound'],3))
PY
  tail -3 gpurun_out/r2b_bench_$c.err >> gpurun_out/r2_call12.log
done
tail -45 gpurun_out/r2_call12.log | cut -c1-330
