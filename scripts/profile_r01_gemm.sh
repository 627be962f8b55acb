set -x
TAG=r01f
STEP="python scripts/probe_step.py --modes graph --steps 205 --reps 1"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 49500 -c 520 --csv --log-file gpurun_out/${TAG}_launches.csv $STEP > gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 39000 -c 5 \
    -o gpurun_out/${TAG}_gemm -f $STEP > gpurun_out/${TAG}_gemm.log 2>&1
tail -3 gpurun_out/${TAG}_gemm.log
