#!/bin/bash
# ncu launch list of one Taming VQGAN decode + encode (16 x 256^2) per precision mode; under gpurun, ONE GPU.
TAG=${1:-r02}
for P in bf16x3 3xtf32; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file gpurun_out/${TAG}_vqgan_${P}_launches.csv python scripts/profile_vqgan.py $P > gpurun_out/${TAG}_vqgan_${P}.log 2>&1
done
ls -la gpurun_out/
