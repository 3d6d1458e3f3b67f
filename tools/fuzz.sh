#!/bin/bash
# more seeds of the randomized parity sweeps (tests/test_gpu_parity.py -k randomized): tools/fuzz.sh [first] [last]
for s in $(seq ${1:-1} ${2:-20}); do
  SRSB200_FUZZ_SEED=$s timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k randomized 2>&1 | tail -1 | sed "s/^/seed $s: /"
done
