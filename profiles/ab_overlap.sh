#!/bin/bash
# A/B of steps in flight x K5 split on one box, alternating (run through gpurun)
B="python bench.py --no-cpu-baseline --no-eager-baseline --steps 100"
for rep in 1 2 3; do
  for cfg in "4 2" "6 1" "6 2" "4 1" "5 1" "8 1"; do
    set -- $cfg
    v=$(KPF_K5_SPLIT=$2 $B --overlap $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'])")
    echo "rep $rep overlap $1 split $2: $v"
  done
done
