#!/bin/bash
# dev helper: effect of the phase-1 sub-batch size (L2 residency of the four-step intermediate)
for u in "$@"; do
  APD_B200_INV_UNITS=$u python bench.py --hours 6 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/benchline.py "inv_units=$u"
done
