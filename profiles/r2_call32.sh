#!/bin/bash
# call 32 (2 GPUs): configs[4] on a 10 % subsample (40 M reads, 13 GB) through bin/sickle on 1 and 2 physical GPUs, file to
# file, input order and -a 16; md5 must not depend on the GPU count; the reference on a prefix.
cd /root/repo
timeout 900 python profiles/multi_gpu_cli.py --reads 40000000 --gpus 1,2 --repeat 1 --ref-reads 4000000 > gpurun_out/r2_call32.log 2>&1
cut -c1-420 gpurun_out/r2_call32.log | tail -8
