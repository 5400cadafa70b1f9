#!/bin/bash
# round 2, GPU call A: full GPU test suite, headline bench (+extras), A/B of the packed-fp32 taps, launch list, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2a_bench.err
SURF_B200_LIB=$PWD/scratch_ab/lib_scalar.so timeout 300 python bench.py --dev-lib --no-extras --no-cpu --steps 20 --warmup 3 > gpurun_out/r2a_bench_scalar_taps.json 2> gpurun_out/r2a_bench_scalar_taps.err; echo "scalar exit $?"
timeout 300 python bench.py --no-extras --no-cpu --steps 20 --warmup 3 > gpurun_out/r2a_bench_packed_taps.json 2> gpurun_out/r2a_bench_packed_taps.err; echo "packed exit $?"
SURF_BENCH_LOOKAHEAD=0 timeout 300 python bench.py --no-extras --no-cpu --steps 20 --warmup 3 > gpurun_out/r2a_bench_nolookahead.json 2>/dev/null
# launch list + captures: the plain command first (must exit 0), then ncu on the same command line
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
if timeout 300 $CMD > gpurun_out/r2a_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2a_launches.csv $CMD > gpurun_out/r2a_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor|k_patch_descriptor' -s 27 -c 9 -o gpurun_out/r2a_describe $CMD > gpurun_out/r2a_ncu_describe.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_candidates|k_hessian_nms|k_band' -s 12 -c 6 -o gpurun_out/r2a_detect_match $CMD > gpurun_out/r2a_ncu_detect.log 2>&1
fi
ls -la gpurun_out | tail -20
