#!/bin/bash
# Round 2, GPU call 20 (8 GPUs): bench.py at N = 1 and N = 8 on the same box (one input sharded over the ranks,
# e2e with the in-run host-link bound), the way the driver launches it.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
L=gpurun_out/r2_call20.log
{
  nvidia-smi topo -m 2>&1 | head -14
  lscpu | grep -E "^CPU\(s\)|NUMA node|Model name|Socket"
  df -h /dev/shm | tail -1; free -g | head -2
} > $L 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8box_n1.json 2> gpurun_out/r2_bench8box_n1.err
echo "== bench N=1 rc=$?" >> $L; tail -c 2600 gpurun_out/r2_bench8box_n1.json >> $L; tail -2 gpurun_out/r2_bench8box_n1.err >> $L
timeout 900 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "== bench N=8 rc=$?" >> $L; tail -c 2600 gpurun_out/r2_bench_n8.json >> $L; tail -3 gpurun_out/r2_bench_n8.err >> $L
tail -c 7000 $L
