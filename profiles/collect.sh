#!/bin/bash
# Collect the per-round profile artifacts on the GPU box (run through gpurun from the repo root):
#   profiles/collect.sh <tag>      -> gpurun_out/<tag>_{launches,dram,stalls}.csv, <tag>_trace.txt, <tag>_bench.json
# All ncu passes are single-pass metric lists: `ncu --set full` hangs on the tcgen05 kernels of this repo.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
common="--clock-control none --profile-from-start off --csv"
KPF_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum $common --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
KPF_PROFILE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum $common --log-file $out/${tag}_dram.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "dram rc=$?"
KPF_PROFILE=1 timeout 300 ncu --metrics regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio,smsp__inst_executed.sum,sm__cycles_elapsed.max \
  $common --log-file $out/${tag}_stalls.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "stalls rc=$?"
KPF_TRACE=$out/${tag}_trace.txt timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "trace rc=$?"
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -1 $out/${tag}_bench.json | cut -c1-400
