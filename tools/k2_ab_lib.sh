#!/bin/bash
# same-box A/B of two library builds: tools/k2_ab_lib.sh <libB.so> [k2_stress args]
B=$1; shift
for i in 1 2 3; do
  echo -n "A  "; python tools/k2_stress.py --reps 10 "$@"
  echo -n "B  "; MVMATCH_LIB_PATH=$B python tools/k2_stress.py --reps 10 "$@"
done
