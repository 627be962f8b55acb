#!/bin/bash
# ncu evidence for the round (run under gpurun; outputs in gpurun_out/, summaries are copied to profiles/ by hand).
# Never a bench value: numbers printed under ncu are discarded.
set -x
TAG=${1:-r01}
STEP="python scripts/probe_step.py --modes graph --steps 205 --reps 1"
# 1) launch list of two token steps around token 200 of the DEFAULT decode path, with DRAM bytes per launch
#    (cold-cache, serialised: compare SHARES only)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 49500 -c 520 --csv --log-file gpurun_out/${TAG}_launches.csv $STEP > gpurun_out/${TAG}_launches.log 2>&1
# 2) --set full capture of the dominant kernel (skinny GEMM, per-GEMM path) and of the fused block kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 39000 -c 5 \
    -o gpurun_out/${TAG}_gemm -f $STEP > gpurun_out/${TAG}_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_block_kernel -s 19200 -c 2 \
    -o gpurun_out/${TAG}_fused -f python scripts/probe_step.py --modes fused --steps 205 --reps 1 > gpurun_out/${TAG}_fused.log 2>&1
ls -la gpurun_out/
