#!/bin/bash
# same-box comparison of kernel 2's cluster modes (0 = single CTA, 2 = multicast pair, 20 = cta_group::2 pair)
for i in 1 2; do
  for c in 0 2 20; do python tools/k2_stress.py --reps 10 --cluster $c "$@"; done
done
