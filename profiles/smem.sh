#!/bin/bash
# shared-memory pipe view of one step (bank conflicts / wavefronts of LSU loads, stores and cp.async writes): profiles/smem.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
KPF_PROFILE=1 timeout 300 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_elapsed.max,sm__cycles_active.sum,smsp__inst_executed.sum \
  --clock-control none --profile-from-start off --csv --log-file $out/${tag}_smem.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --overlap 1 > /dev/null 2>&1; echo "smem rc=$?"
