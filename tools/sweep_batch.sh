#!/bin/bash
# dev helper: tools/sweep_batch.sh "<batch> [ENV=..]" ...  -> one short bench line per (batch size, environment)
HOURS=${HOURS:-24}
WL=${WL:-c3}
for spec in "$@"; do
  set -- $spec
  b=$1; shift
  env "$@" python bench.py --workload $WL --hours $HOURS --steps 3 --warmup 2 --batch-chunks $b --no-cpu-baseline 2>/dev/null \
    | python tools/benchline.py "[batch=$b $*]"
done
