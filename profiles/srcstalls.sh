#!/bin/bash
# per-source-line warp-stall samples of one kernel: profiles/srcstalls.sh <kernel regex> <tag> [driver script]   (run through gpurun)
k=${1:-desa_tile}
tag=${2:-r2}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k regex:$k -s 6 -c 1 \
  -o $out/${tag}_src -f python ${3:-profiles/probe_kernels.py} > $out/${tag}_src.log 2>&1; echo "ncu rc=$?"
ncu -i $out/${tag}_src.ncu-rep --page source --csv > $out/${tag}_src.csv 2> $out/${tag}_src.err; echo "export rc=$?"
