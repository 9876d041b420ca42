#!/bin/bash
# GPU soak: the seeded random parity tests (single end x quality types x lengths up to 3 kb, pairs in four modes,
# -a N order, damaged inputs; both kernel paths) drawn with other seeds, results against the oracle.
#   usage (on the GPU box, from the repo root): profiles/gpu_soak.sh [first_seed] [n_seeds]
cd "$(dirname "$0")/.."
FIRST=${1:-1}; N=${2:-10}
pass=0; fail=0
for s in $(seq $FIRST $((FIRST + N - 1))); do
  if SICKLE_B200_SOAK_SEED=$s timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x \
       -k "random_se_vs_oracle or random_pe_vs_oracle or emulated_thread_order or fuzzed_inputs" > /tmp/soak_$s.log 2>&1; then
    pass=$((pass + 1)); echo "seed $s: $(tail -1 /tmp/soak_$s.log)"
  else
    fail=$((fail + 1)); echo "seed $s: FAILED"; tail -30 /tmp/soak_$s.log
  fi
done
echo "soak: $pass seeds passed, $fail failed"
