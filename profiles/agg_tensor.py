"""Per-kernel tensor-pipe activity and memory-pipe request counts from profiles/collect.sh's `<tag>_tensor.csv` (single-pass ncu list)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "").replace("kpf::", "").strip()
    try:
        agg[name][row["Metric Name"]].append(float(row["Metric Value"].replace(",", "")))
    except ValueError:
        pass
cols = [("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of elapsed"), ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "LDG requests"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "LDG sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ldgsts.sum", "LDGSTS req"), ("lts__t_sectors_op_read.sum", "L2 rd sectors"),
        ("lts__t_sectors_op_write.sum", "L2 wr sectors"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs")]
print("per launch (mean over the profiled launches), B = 64, one B200, `--clock-control none`")
print(f"{'kernel':30s} " + " ".join(f"{c[1]:>24s}" for c in cols))
for k, m in agg.items():
    print(f"{k[:30]:30s} " + " ".join(f"{(sum(m[c[0]]) / len(m[c[0]])) if m.get(c[0]) else float('nan'):24.2f}" for c in cols))
