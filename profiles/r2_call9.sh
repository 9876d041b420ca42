#!/bin/bash
# Round 2, GPU call 9: (a) what the phases cost in paired-end mode (knock-outs, phase timing);
# (b) the command line file to file: start-up split, write(2) vs mapped output (with and without fallocate), tmpfs ceilings.
mkdir -p gpurun_out
V=build/variants
S=sickle_b200/libsickle_b200.so
{
  echo "== pe interleaved: knock-outs"
  python profiles/ab_multi.py --workload pe $S $V/lib_KO_S6.so $V/lib_KO_S8A.so $V/lib_KO_FLUSH.so $V/lib_KO_LB1.so $V/lib_KO_LB2.so
  echo "== pe -M: knock-outs"
  python profiles/ab_multi.py --workload pem $S $V/lib_KO_FLUSH.so $V/lib_KO_LB2.so
  echo "== se: knock-outs (current library)"
  python profiles/ab_multi.py $S $V/lib_KO_S6.so $V/lib_KO_S8A.so $V/lib_KO_FLUSH.so $V/lib_KO_LB1.so $V/lib_KO_LB2.so
  echo "== io_probe"
  bin/io_probe /dev/shm 2048
  echo "== cli, write(2)"
  SICKLE_B200_DEBUG_INIT=1 python profiles/cli_bench.py --reads 24000000 --skip-ref --repeat 2
  echo "== cli, mapped output (ftruncate)"
  python profiles/cli_bench.py --reads 24000000 --skip-ref --repeat 2 --env SICKLE_B200_MMAP_OUT=1
  echo "== cli, mapped output (fallocate first)"
  python profiles/cli_bench.py --reads 24000000 --skip-ref --repeat 2 --env SICKLE_B200_MMAP_OUT=2
} > gpurun_out/r2_call9.log 2>&1
tail -60 gpurun_out/r2_call9.log | cut -c1-700
