#!/bin/bash
# Round 2, final single-GPU call: the whole GPU test suite, smoke, bench.py (both arms, every config), the ncu launch
# list of the bench command, `--set full` captures of every kernel, ncu_traffic.json, workloads, the command line.
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
{
  nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
  echo "== pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
  echo "== smoke"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > $L 2>&1
# ---- ncu: launch list of the bench command, then full sets
# (our kernels only: the synthetic input alone is thousands of torch launches)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kf_fused|kf_finalize|kf2_|kfo_|k1_line_index|k2_trim|k3_emit|k_finalize" -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 4 --warmup 3 --kernel-only --min-timed-s 0.001 > gpurun_out/r2_final_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kf_fused|kf_finalize|kf2_|kfo_|k1_line_index|k2_trim|k3_emit|k_finalize" -c 400 --csv --log-file gpurun_out/r2_final_launches_c3.csv python bench.py --config c3 --steps 4 --warmup 3 --kernel-only --min-timed-s 0.001 >> gpurun_out/r2_final_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kf_fused|kf_finalize|kf2_|kfo_|k1_line_index|k2_trim|k3_emit|k_finalize" -c 400 --csv --log-file gpurun_out/r2_final_launches_c4.csv python bench.py --config c4 --steps 4 --warmup 3 --kernel-only --min-timed-s 0.001 >> gpurun_out/r2_final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kf_fused -s 4 -c 1 -o gpurun_out/r2_final_fused -f python bench.py --steps 6 --warmup 3 --kernel-only --min-timed-s 0.001 > gpurun_out/r2_final_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kf_fused -s 8 -c 2 -o gpurun_out/r2_final_twofile -f python bench.py --config c3 --steps 6 --warmup 3 --kernel-only --min-timed-s 0.001 > gpurun_out/r2_final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_line_index|k2_trim_route|k3_emit" -s 6 -c 3 -o gpurun_out/r2_final_general -f python profiles/workloads.py --general-only > gpurun_out/r2_final_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_line_index|k2_trim|k3_emit" -s 12 -c 4 -o gpurun_out/r2_final_long -f python profiles/workloads.py --c4-only -x -n > gpurun_out/r2_final_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kf_fused|kfo_" -s 8 -c 4 -o gpurun_out/r2_final_ordered -f python profiles/workloads.py --a8-only > gpurun_out/r2_final_ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kf_fused|kf_finalize|kf2_|kfo_|k1_line_index|k2_trim|k3_emit|k_finalize" -c 60 --csv --log-file gpurun_out/r2_final_launches_a8.csv python profiles/workloads.py --a8-only >> gpurun_out/r2_final_ncu_launches.log 2>&1
for r in fused twofile general long ordered; do
  ncu -i gpurun_out/r2_final_$r.ncu-rep --page raw --csv > gpurun_out/r2_final_${r}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_final_fused.ncu-rep --page source --csv > gpurun_out/r2_final_fused_source.csv 2>/dev/null
python profiles/make_ncu_traffic.py c2=gpurun_out/r2_final_fused_raw.csv c3=gpurun_out/r2_final_twofile_raw.csv c4=gpurun_out/r2_final_long_raw.csv > gpurun_out/r2_final_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
python bench.py > gpurun_out/r2_final_bench_c2_default.json 2> gpurun_out/r2_final_bench_c2_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
for c in c2 c3 c3m c4; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_$c.json 2> gpurun_out/r2_final_bench_$c.err
done
python - >> $L 2>&1 <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_final_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    if d.get('impl') == 'reference':
        print(f, 'reference', round(d['value']), 'reads/s', d['cpu_baseline']); continue
    e = d.get('e2e') or {}
    print(f, 'ms/step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'traffic', d['roofline']['traffic'], d['roofline']['stage_ms'],
          'e2e', e and round(e['value'] / 1e6, 1), e and round(e['frac_of_bound'], 3), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'clocks', d['clocks'])
PY
{
  echo "== ncu_traffic.json"; cat profiles/ncu_traffic.json | head -40
  echo "== workloads"
  python profiles/workloads.py
  echo "== command line, 24 M reads"
  python profiles/cli_bench.py --reads 24000000 --skip-ref --repeat 2
  python profiles/cli_bench.py --reads 8000000 --repeat 1
  echo "== tile sizes, single end (same process)"
  python profiles/ab_multi.py --workload se --rounds 3 --steps 24 sickle_b200/libsickle_b200.so@SICKLE_B200_FUSED_CH=7 sickle_b200/libsickle_b200.so@SICKLE_B200_FUSED_CH=9 sickle_b200/libsickle_b200.so@SICKLE_B200_FUSED_CH=11
} >> $L 2>&1
tail -60 $L | cut -c1-420
