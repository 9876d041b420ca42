#!/bin/bash
# Round 2, GPU call 6: where K2 spends its time on configs[3] (flag combinations, ncu), then the sanitizers.
mkdir -p gpurun_out
{
  echo "== configs[3] K2 by flags"
  python profiles/workloads.py --c4-only
  python profiles/workloads.py --c4-only -x
  python profiles/workloads.py --c4-only -n
  python profiles/workloads.py --c4-only -x -n
  echo "== sanitize_run without a tool"
  timeout 300 python profiles/sanitize_run.py 2>&1 | tail -25
} > gpurun_out/r2_call6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_trim_route -s 3 -c 1 -o gpurun_out/r2_k2_c4 -f python profiles/workloads.py --c4-only -x -n > gpurun_out/r2_call6_ncu.log 2>&1
ncu -i gpurun_out/r2_k2_c4.ncu-rep --page raw --csv > gpurun_out/r2_k2_c4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_k2_c4.ncu-rep --page source --csv > gpurun_out/r2_k2_c4_source.csv 2>/dev/null
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== compute-sanitizer $tool: rc=$? $(grep -c SANITIZE_RUN_OK gpurun_out/r2_sanitizer_$tool.log) ok-lines; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)" >> gpurun_out/r2_call6.log
done
tail -45 gpurun_out/r2_call6.log
