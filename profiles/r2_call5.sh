#!/bin/bash
# Round 2, GPU call 5: configs[0] fixtures on the CUDA path; workloads after the long-read work; compute-sanitizer
# (memcheck, racecheck, synccheck, initcheck) over every kernel with results checked against the oracle.
mkdir -p gpurun_out
{
  echo "== reference fixtures (configs[0]) through the C ABI and the CLI"
  timeout 600 python -m pytest tests/test_reference_fixtures.py -m gpu -x -q 2>&1 | tail -4
  echo "== workloads"
  python profiles/workloads.py
  echo "== sanitize_run without a tool"
  timeout 300 python profiles/sanitize_run.py 2>&1 | tail -25
} > gpurun_out/r2_call5.log 2>&1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== compute-sanitizer $tool: rc=$? $(grep -c SANITIZE_RUN_OK gpurun_out/r2_sanitizer_$tool.log) ok-lines; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)" >> gpurun_out/r2_call5.log
done
tail -45 gpurun_out/r2_call5.log
