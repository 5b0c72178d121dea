#!/bin/bash
# dev helper: tools/sweep_alone.sh "A=1 B=2" "A=3" ...   -> correlate stage alone per environment
HOURS=${HOURS:-6}
for e in "$@"; do
  echo -n "[$e] "; env $e python tools/corr_alone.py $HOURS 2>&1 | tail -1
done
