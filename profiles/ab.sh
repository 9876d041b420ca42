#!/bin/bash
# A/B two builds of the library on the same GPU: usage profiles/ab.sh libA.so libB.so [rounds]
for r in $(seq 1 ${3:-3}); do
  for lib in "$1" "$2"; do
    SICKLE_B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 3 --kernel-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
