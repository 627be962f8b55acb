#!/bin/bash
# ncu evidence for the round (run under gpurun; outputs in gpurun_out/, summaries are copied to profiles/ by hand).
# 1) launch list of two token steps deep inside the timed region (cold-cache, serialised: compare SHARES only)
# 2) --set full capture of the dominant kernels
set -x
TAG=${1:-r01}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 111000 -c 520 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 40000 -c 5 \
    -o gpurun_out/${TAG}_gemm -f $BENCH > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 3 \
    -o gpurun_out/${TAG}_conv -f $BENCH > gpurun_out/${TAG}_conv.log 2>&1
ls -la gpurun_out/
