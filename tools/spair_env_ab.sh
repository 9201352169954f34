#!/bin/bash
# same-box sweep of the streaming SPair kernel's run-time switches (MVMATCH_SPAIR_<KEY>): tools/spair_env_ab.sh <pairs> "STAGES=2" "MMA=0" "STREAM=0" ...
# (chunk size, CTAs per SM, L2 policy, Q warps are compile-time: -DSPS_CC_N / SPS_CTAS / SPS_HINTS / SPS_QREV / SPS_QWARPS, see tools/spair_ab.sh)
P=$1; shift
for r in 1 2; do
  for V in "$@"; do
    echo -n "$V  "
    env $(for kv in $V; do echo -n "MVMATCH_SPAIR_$kv "; done) python tools/spair_probe.py --pairs $P 2>/dev/null | head -1
  done
done
