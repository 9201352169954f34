#!/bin/bash
# same-box A/B of SPair-kernel builds: tools/spair_ab.sh <pairs> lib1.so lib2.so ...   ("default" = the in-tree build)
P=$1; shift
for r in 1 2; do
  for L in "$@"; do
    echo -n "$(basename $L)  "
    if [ "$L" = default ]; then python tools/spair_probe.py --pairs $P 2>/dev/null | head -1
    else MVMATCH_LIB_PATH=$L python tools/spair_probe.py --pairs $P 2>/dev/null | head -1; fi
  done
done
