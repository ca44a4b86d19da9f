#!/bin/bash
# Collect the per-round profile artifacts on the GPU box (run through gpurun from the repo root):
#   profiles/collect.sh <tag>      -> gpurun_out/<tag>_{launches,dram,stalls,tensor}.csv, <tag>_trace.txt, <tag>_bench.json
# All ncu passes are single-pass metric lists around two warm steps (KPF_PROFILE=1 -> cudaProfilerStart/Stop in bench.py):
# `ncu --set full` hangs on the tcgen05 kernels of this repo (round 1), and a number printed by a run under ncu is never a bench value.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
common="--clock-control none --profile-from-start off --csv"
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline"
KPF_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum $common --log-file $out/${tag}_launches.csv $cmd > /dev/null 2>&1; echo "launches rc=$?"
KPF_PROFILE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum $common --log-file $out/${tag}_dram.csv $cmd > /dev/null 2>&1; echo "dram rc=$?"
KPF_PROFILE=1 timeout 300 ncu --metrics regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio,smsp__inst_executed.sum,sm__cycles_elapsed.max \
  $common --log-file $out/${tag}_stalls.csv $cmd > /dev/null 2>&1; echo "stalls rc=$?"
# tensor-pipe activity (north_star: "tensor-pipe utilisation reported against the bf16 dense peak"), memory-pipe requests, occupancy
KPF_PROFILE=1 timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.sum,sm__inst_executed_pipe_tensor.sum,sm__cycles_active.sum,sm__cycles_elapsed.max,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ldgsts.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  $common --log-file $out/${tag}_tensor.csv $cmd > /dev/null 2>&1; echo "tensor rc=$?"
KPF_TRACE=$out/${tag}_trace.txt timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > /dev/null 2>&1; echo "trace rc=$?"
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -1 $out/${tag}_bench.json | cut -c1-400
