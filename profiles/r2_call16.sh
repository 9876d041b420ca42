#!/bin/bash
# Round 2, GPU call 16: look-back #1 by one warp, its loads in flight under the position stores (A/B against the 8-warp walk).
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== parity"
  timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
  echo "== se"
  python profiles/ab_multi.py $V/lib_r1.so $V/lib_prev3.so $S
  echo "== se CH=7"
  SICKLE_B200_FUSED_CH=7 python profiles/ab_multi.py $V/lib_prev3.so $S
  echo "== pe interleaved"
  python profiles/ab_multi.py --workload pe $V/lib_prev3.so $S
} > gpurun_out/r2_call16.log 2>&1
timeout 600 python bench.py --config c3 --steps 20 --warmup 3 --kernel-only > gpurun_out/r2_call16_c3.json 2>/dev/null
python - >> gpurun_out/r2_call16.log <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call16_c3.json').read().strip().splitlines()[-1])
print('bench c3 ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
PY
tail -20 gpurun_out/r2_call16.log | cut -c1-250
