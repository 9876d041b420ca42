#!/bin/bash
# Round 2, GPU call 15: K2 as two kernels for batches of long records, K3 with one record per warp (configs[3]).
mkdir -p gpurun_out
{
  echo "== parity"
  timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_reference_fixtures.py -m gpu -x -q 2>&1 | tail -3
  echo "== workloads"
  python profiles/workloads.py
} > gpurun_out/r2_call15.log 2>&1
timeout 600 python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_call15_bench_c4.json 2> gpurun_out/r2_call15_bench_c4.err
python - >> gpurun_out/r2_call15.log <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call15_bench_c4.json').read().strip().splitlines()[-1])
print('bench c4 ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), d['roofline']['stage_ms'], 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['value']/1e6,2))
PY
tail -25 gpurun_out/r2_call15.log | cut -c1-330
