#!/bin/bash
# Per-kernel warp stall reasons of two warm steps (single-pass counters only: `ncu --set full` hangs on the tcgen05 kernels).
# usage: profiles/stalls.sh <out.csv> [kernel regex]
out=${1:-gpurun_out/stalls.csv}
k=${2:-"token_stack|desa_fused|point_embed|spatial_aggregate_tc|nearest"}
KPF_PROFILE=1 timeout 200 ncu --metrics regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio,smsp__inst_executed.sum,sm__cycles_elapsed.max \
  --clock-control none --profile-from-start off -k regex:"$k" -c 16 --csv --log-file "$out" \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "rc=$? $(wc -l < "$out") lines"
