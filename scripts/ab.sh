#!/bin/bash
# A/B of engine libraries on one box: scripts/ab.sh label=path[,ENV=VAL] ...   (em_ms and resident sites/s per variant, two rounds)
for round in 1 2; do
for spec in "$@"; do
  label=${spec%%=*}; rest=${spec#*=}
  lib=${rest%%,*}; envs=""
  if [[ "$rest" == *,* ]]; then envs=${rest#*,}; fi
  ( export CRISPGPU_LIB=$lib; IFS=,; for e in $envs; do export $e; done
    python bench.py --steps 20 --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$label', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'em_ms %.4f ev_ms %.4f perm_ms %.4f' % (k['em_ms'], k['evaluate_ms'], k['permute_ms']))" )
done; done
