#!/bin/bash
# First GPU call of the next round, everything that was prepared on CPU and is waiting for a B200:
#   gpurun --timeout 1500 -- 'bash profiles/next_round.sh'
# Results land in gpurun_out/nr_*.  Every timing below is taken without a profiler.
mkdir -p gpurun_out
{
  echo "== new GPU tests (several-device CLI, config 4, parallel golden CLI)"
  timeout 900 python -m pytest tests/test_cuda_parity.py::test_long_reads_config4 tests/test_cli.py -m gpu -x -q --durations=8 2>&1 | tail -15
  echo "== bench (shipped library)"
  python bench.py 2>/dev/null | tail -1
  echo "== build variants: parity + same-GPU A/B"
  bash profiles/ab_variants.sh 3
  echo "== the direct-emit variants again, both libraries pinned to the 25 KB tile (3 CTAs per SM with two input buffers)"
  SICKLE_B200_FUSED_CH=7 bash profiles/ab_variants.sh 3 SK_DIRECT_EMIT "SK_DIRECT_EMIT -DSK_EARLY_LOAD"
  echo "== knock-outs: what each phase costs in throughput (output wrong by construction)"
  bash profiles/ab_variants.sh 2 SK_KO_S6 SK_KO_S8A SK_KO_FLUSH SK_KO_LB1 SK_KO_LB2
  echo "== other shapes of the path, incl. BASELINE configs[3]"
  python profiles/workloads.py
  echo "== the program, one context vs two contexts on this GPU (file to file on tmpfs)"
  python profiles/cli_bench.py --reads 8000000 --skip-ref --repeat 2
  python profiles/cli_bench.py --reads 8000000 --skip-ref --repeat 2 --env SICKLE_B200_DEVICES=0,0
} > gpurun_out/nr_first_call.log 2>&1
tail -60 gpurun_out/nr_first_call.log
