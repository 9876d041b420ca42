#!/bin/bash
# Same-GPU A/B of the experimental build variants against the shipped library.
#   usage (on the GPU box, from the repo root):  profiles/ab_variants.sh [rounds] [VARIANT ...]
# For every variant (default: all of the list below, plus their combination): build it with
# `make variant V=<macro>`, run the CUDA parity tests against it (SICKLE_B200_LIB), then alternate
# kernel-only bench runs of the shipped library and the variant `rounds` times (default 3) and print
# ms per 1M-read step and the roofline fraction of each run.  A variant that fails parity is not timed.
# `profiles/ab_variants.sh 2 SK_KO_S6 SK_KO_S8A SK_KO_FLUSH SK_KO_LB1 SK_KO_LB2` times the knock-outs
# (one phase removed each, output wrong by construction): what each phase costs in throughput.
set -u
ROUNDS=${1:-3}
shift || true
VARIANTS=("$@")
[ ${#VARIANTS[@]} -eq 0 ] && VARIANTS=(SK_LANE_SPLIT4 SK_STAGE_LONG_FIRST SK_NL_BRANCHFREE "SK_LANE_SPLIT4 -DSK_STAGE_LONG_FIRST -DSK_NL_BRANCHFREE" SK_DIRECT_EMIT "SK_DIRECT_EMIT -DSK_EARLY_LOAD")
# (SK_DIRECT_EMIT doubles the input buffer: 2 CTAs per SM at the default 32 KB tile, 3 at 25 KB -- also run it as
#  SICKLE_B200_FUSED_CH=7 profiles/ab_variants.sh 3 SK_DIRECT_EMIT, which pins both libraries to the 25 KB tile)
BASE=sickle_b200/libsickle_b200.so
one() {   # lib label
  SICKLE_B200_LIB=$PWD/$1 python bench.py --steps 30 --warmup 3 --kernel-only 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
}
for V in "${VARIANTS[@]}"; do
  NAME=$(echo "$V" | tr -d ' ' | sed 's/-D/+/g')
  LIB=sickle_b200/libsickle_b200_${NAME}.so
  echo "=== $V"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
      -D$V -shared sickle_b200/csrc/capi.cu -o $LIB 2> /dev/null || { echo "build failed"; continue; }
  case "$V" in
    SK_KO_*) echo "(timing-only knock-out: output is wrong by construction, no parity run)";;
    *) env -u SICKLE_B200_FUSED_CH SICKLE_B200_LIB=$PWD/$LIB python -m pytest tests/test_cuda_parity.py -m gpu -x -q > /tmp/ab_parity.log 2>&1   # (a pinned tile size is for the timing runs only)
       RC=$?
       tail -2 /tmp/ab_parity.log
       [ $RC -ne 0 ] && { echo "PARITY FAILED: not timed"; continue; };;
  esac
  for r in $(seq 1 $ROUNDS); do one $BASE shipped; one $LIB "$NAME"; done
done
