#!/bin/bash
# Round 2, GPU call 8 (4 GPUs): the product on physical GPUs -- several-device CLI tests, one file through
# bin/sickle on 1 / 2 / 4 GPUs (md5 must not depend on the GPU count), bench.py at N = 2 and 4 (one input sharded
# over the ranks, e2e with the in-run host-link bound), affinity and write-combined A/B at N = 4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  nvidia-smi topo -m 2>&1 | head -20
  lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" 
  echo "== several-device CLI tests on physical GPUs"
  timeout 900 python -m pytest tests/test_cli.py -m gpu -k "several_devices" -q 2>&1 | tail -4
  echo "== one file, 1 / 2 / 4 GPUs"
  timeout 1200 python profiles/multi_gpu_cli.py --reads 24000000 --gpus 1,2,4
} > gpurun_out/r2_call8.log 2>&1
for n in 2 4; do
  timeout 900 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "== bench N=$n rc=$?" >> gpurun_out/r2_call8.log; tail -c 2500 gpurun_out/r2_bench_n$n.json >> gpurun_out/r2_call8.log; tail -3 gpurun_out/r2_bench_n$n.err >> gpurun_out/r2_call8.log
done
BENCH_NO_AFFINITY=1 timeout 900 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4_noaff.json 2> gpurun_out/r2_bench_n4_noaff.err
echo "== bench N=4 no affinity rc=$?" >> gpurun_out/r2_call8.log; tail -c 2000 gpurun_out/r2_bench_n4_noaff.json >> gpurun_out/r2_call8.log
SICKLE_B200_WC_INPUT=1 timeout 900 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4_wc.json 2> gpurun_out/r2_bench_n4_wc.err
echo "== bench N=4 write-combined input rc=$?" >> gpurun_out/r2_call8.log; tail -c 2000 gpurun_out/r2_bench_n4_wc.json >> gpurun_out/r2_call8.log
tail -c 6000 gpurun_out/r2_call8.log
