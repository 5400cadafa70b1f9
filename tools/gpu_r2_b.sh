#!/bin/bash
# round 2, GPU call B: full GPU suite (per-test timeout), headline bench (+extras), A/B of the pipelined matcher, launch list, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=240 -rf -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest_gpu.log
tail -15 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2b_bench.err
SURF_B200_LIB=$PWD/scratch_ab/lib_tcserial.so timeout 300 python bench.py --dev-lib --no-extras --no-cpu --steps 20 --warmup 3 > gpurun_out/r2b_bench_tcserial.json 2> gpurun_out/r2b_bench_tcserial.err; echo "tcserial exit $?"
timeout 300 python bench.py --no-extras --no-cpu --steps 100 --warmup 3 > gpurun_out/r2b_bench_100steps.json 2> gpurun_out/r2b_bench_100steps.err; echo "100 steps exit $?"
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
if timeout 300 $CMD > gpurun_out/r2b_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2b_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor' -s 15 -c 5 -o gpurun_out/r2b_describe $CMD > gpurun_out/r2b_ncu_describe.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_candidates|k_hessian_nms|k_band' -s 12 -c 6 -o gpurun_out/r2b_detect_match $CMD > gpurun_out/r2b_ncu_detect.log 2>&1
fi
ls -la gpurun_out | grep r2b
