#!/bin/bash
# round 2, GPU call I2: ncu --set full of the describe stage, the fused Hessian kernels and the integral kernel at HEAD (one step), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 300 $CMD > gpurun_out/r2i2_bench.json 2>/dev/null; echo "bench exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor' -s 15 -c 5 -o gpurun_out/r2i2_describe $CMD > gpurun_out/r2i2_ncu_describe.log 2>&1; echo "ncu describe exit $?"
timeout 600 ncu --set full --clock-control none -k regex:'k_hessian_nms|k_integral_band|k_sort_chunks|k_pack' -s 12 -c 5 -o gpurun_out/r2i2_detect $CMD > gpurun_out/r2i2_ncu_detect.log 2>&1; echo "ncu detect exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2i2_launches.csv $CMD > /dev/null 2>&1; echo "ncu launches exit $?"
