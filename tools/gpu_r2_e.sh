#!/bin/bash
# round 2, GPU call E: parity after the streaming big-window kernel, bench, e2e policy A/B on one box, describe-stage ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=240 -rf -p no:cacheprovider > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest_gpu.log
tail -8 gpurun_out/r2e_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
SURF_B200_EARLY_UPLOAD=0 timeout 300 $B > gpurun_out/r2e_bench_late.json 2> gpurun_out/r2e_bench_late.err; echo "late exit $?"
SURF_B200_EARLY_UPLOAD=1 timeout 300 $B > gpurun_out/r2e_bench_early.json 2>/dev/null; echo "early exit $?"
SURF_B200_LIB=$PWD/scratch_ab/lib_callA.so timeout 300 $B --dev-lib > gpurun_out/r2e_bench_callA.json 2>/dev/null; echo "callA exit $?"
SURF_B200_EARLY_UPLOAD=0 timeout 300 $B > gpurun_out/r2e_bench_late2.json 2>/dev/null; echo "late2 exit $?"
SURF_B200_EARLY_UPLOAD=1 timeout 300 $B > gpurun_out/r2e_bench_early2.json 2>/dev/null; echo "early2 exit $?"
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
if timeout 300 $CMD > gpurun_out/r2e_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2e_launches.csv $CMD > gpurun_out/r2e_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor' -s 15 -c 5 -o gpurun_out/r2e_describe $CMD > gpurun_out/r2e_ncu_describe.log 2>&1
fi
ls -la gpurun_out | grep r2e
