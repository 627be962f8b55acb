#!/bin/bash
# ncu evidence of the round-2 final state (run under gpurun, ONE GPU; outputs in gpurun_out/, summaries copied to profiles/).
# Never a bench value: numbers printed under ncu are discarded.
set -x
TAG=${1:-r02_final}
STEP="python scripts/probe_step.py --modes graph --steps 205 --reps 1"
# (1) launch list of two token steps around token 200 of the default Taming path (shares only: cold cache, serialised)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 48400 -c 500 --csv --log-file gpurun_out/${TAG}_launches_token200.csv $STEP > gpurun_out/${TAG}_launches.log 2>&1
# (2) --set full on the skinny GEMM (TMA weight ring) around token 200: one layer's four GEMMs
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm_kernel -s 38400 -c 4 \
    -o gpurun_out/${TAG}_gemm -f $STEP > gpurun_out/${TAG}_gemm.log 2>&1
ncu -i gpurun_out/${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_ncu_raw.csv 2>/dev/null
# (3) --set full on the persistent bf16x3 tcgen05 conv: the 256^2 128->128 and the 64^2 256->256 layers of one decode
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_bf16 -s 60 -c 18 \
    -o gpurun_out/${TAG}_convbf16 -f python scripts/profile_vqgan.py bf16x3 > gpurun_out/${TAG}_convbf16.log 2>&1
ncu -i gpurun_out/${TAG}_convbf16.ncu-rep --page raw --csv > gpurun_out/${TAG}_convbf16_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
