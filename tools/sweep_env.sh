#!/bin/bash
# dev helper: tools/sweep_env.sh "A=1 B=2" "A=3" ...   -> one short bench line per environment
HOURS=${HOURS:-6}
for e in "$@"; do
  env $e python bench.py --hours $HOURS --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/benchline.py "[$e]"
done
