#!/bin/bash
# A/B of programmatic dependent launch with six steps in flight (run through gpurun)
B="python bench.py --no-cpu-baseline --no-eager-baseline --steps 100"
for rep in 1 2 3; do
  for pdl in 1 0; do
    v=$(KPF_PDL=$pdl $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'])")
    echo "rep $rep KPF_PDL=$pdl: $v"
  done
done
