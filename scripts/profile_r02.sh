#!/bin/bash
# ncu evidence planned for round 2 (run under gpurun, ONE GPU; outputs in gpurun_out/, summaries go to profiles/ by hand).
# Never a bench value: numbers printed under ncu are discarded.  Answers the open questions of DESIGN.md section 8:
#   a) where the attention kernel's time goes at token 200 (source-level warp states),
#   b) why the activation loads of skinny_gemm_bf16_kernel cost 19 % of an Anole-7B pass (knock-out) although neither
#      their latency nor their bytes matter.
set -x
TAG=${1:-r02}
STEP="python scripts/probe_step.py --modes graph --steps 205 --reps 1"
# launch list of two token steps around token 200 of the default Taming path (shares only: cold cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 49500 -c 520 --csv --log-file gpurun_out/${TAG}_launches.csv $STEP > gpurun_out/${TAG}_launches.log 2>&1
# a) attention kernel, two launches near token 200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 9600 -c 2 \
    -o gpurun_out/${TAG}_attn -f $STEP > gpurun_out/${TAG}_attn.log 2>&1
# b) bf16 GEMM on Anole-7B shapes: one full layer (wqkv, wo, w13, w2) of a late pass
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm_bf16 -s 2560 -c 4 \
    -o gpurun_out/${TAG}_bf16 -f python scripts/probe_cham.py 5 32 > gpurun_out/${TAG}_bf16.log 2>&1
ls -la gpurun_out/
# read here with:
#   ncu -i gpurun_out/${TAG}_bf16.ncu-rep --page source --csv --kernel-name regex:skinny --launch-skip 2 --launch-count 1
#   ncu -i gpurun_out/${TAG}_attn.ncu-rep --page raw --csv
