#!/bin/bash
# round 2, GPU call C: parity of the describe changes, A/B builds (call-A library, packed row pass), e2e diagnosis
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q --timeout=240 -rf -p no:cacheprovider > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log
tail -5 gpurun_out/r2c_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
timeout 300 $B > gpurun_out/r2c_bench_head.json 2> gpurun_out/r2c_bench_head.err; echo "head exit $?"
SURF_B200_LIB=$PWD/scratch_ab/lib_callA.so timeout 300 $B --dev-lib > gpurun_out/r2c_bench_callA.json 2> gpurun_out/r2c_bench_callA.err; echo "callA exit $?"
SURF_B200_LIB=$PWD/scratch_ab/lib_rowpacked.so timeout 300 $B --dev-lib > gpurun_out/r2c_bench_rowpacked.json 2> gpurun_out/r2c_bench_rowpacked.err; echo "rowpacked exit $?"
SURF_BENCH_E2E_DIAG=no_d2h timeout 300 $B > gpurun_out/r2c_bench_e2e_no_d2h.json 2>/dev/null; echo "no_d2h exit $?"
SURF_BENCH_E2E_DIAG=no_h2d timeout 300 $B > gpurun_out/r2c_bench_e2e_no_h2d.json 2>/dev/null; echo "no_h2d exit $?"
timeout 300 python bench.py --no-extras --no-cpu --steps 150 --warmup 3 > gpurun_out/r2c_bench_head_150.json 2>/dev/null; echo "150 exit $?"
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
if timeout 300 $CMD > gpurun_out/r2c_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor' -s 15 -c 5 -o gpurun_out/r2c_describe $CMD > gpurun_out/r2c_ncu_describe.log 2>&1
fi
ls -la gpurun_out | grep r2c
