#!/bin/bash
# same-box sweep of the streaming SPair kernel's launch plan: tools/spair_env_ab.sh <pairs> "CC=24 PAD=1" "CC=16 STAGES=3" ...
P=$1; shift
for r in 1 2; do
  for V in "$@"; do
    echo -n "$V  "
    env $(for kv in $V; do echo -n "MVMATCH_SPAIR_$kv "; done) python tools/spair_probe.py --pairs $P 2>/dev/null | head -1
  done
done
