#!/bin/bash
# A/B two settings of MVMATCH_K2_FLAGS on the same box, interleaved (power / clock state drifts between runs)
for i in 1 2 3; do
  for f in 0 1; do
    echo -n "flags=$f  "; MVMATCH_K2_FLAGS=$f python tools/k2_stress.py --reps 10 "$@"
  done
done
