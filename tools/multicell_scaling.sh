#!/bin/bash
# BASELINE config 5 (64 cells x 13 code blocks per subframe) from plain C on 1..N GPUs of one box, one process:
# cells are placed on GPUs by cell id, no collective. Usage (on the GPU box): tools/multicell_scaling.sh <max gpus> > out.jsonl
max=${1:-8}
nproc >&2
for g in 1 2 4 8; do
  [ $g -le $max ] || break
  for t in 2 4 8; do
    [ $((g * t)) -le 64 ] || continue
    for rep in 1 2; do
      timeout 120 examples/multicell_uplink $t 64 50 1 $g
    done
  done
done
