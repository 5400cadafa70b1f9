#!/bin/bash
# round 2, GPU call F: parity + bench after the row-sum pass changes, then compute-sanitizer memcheck on the small run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q --timeout=240 -rf -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest_gpu.log
tail -4 gpurun_out/r2f_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
timeout 300 $B > gpurun_out/r2f_bench_head.json 2> gpurun_out/r2f_bench_head.err; echo "head exit $?"
SURF_B200_LIB=$PWD/scratch_ab/lib_callA.so timeout 300 $B --dev-lib > gpurun_out/r2f_bench_callA.json 2>/dev/null; echo "callA exit $?"
timeout 300 python tools/sanitize_small.py > gpurun_out/r2f_sanitize_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck --print-limit 50 python tools/sanitize_small.py > gpurun_out/r2f_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/r2f_memcheck.log
ls -la gpurun_out | grep r2f
