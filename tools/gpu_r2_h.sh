#!/bin/bash
# round 2, GPU call H (1 GPU): full GPU suite, smoke, the default bench run (wall-clocked) and the reference arm at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2h_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2h_pytest_gpu.log
tail -4 gpurun_out/r2h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke exit $?"
S=$(date +%s.%N)
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench exit $?"
E=$(date +%s.%N); echo "bench default wall $(echo "$E - $S" | bc) s" | tee gpurun_out/r2h_bench_wall.txt
S=$(date +%s.%N)
timeout 900 python bench.py --impl reference > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err; echo "ref exit $?"
E=$(date +%s.%N); echo "bench reference wall $(echo "$E - $S" | bc) s" | tee -a gpurun_out/r2h_bench_wall.txt
head -c 700 gpurun_out/r2h_bench_default.json; echo
head -c 500 gpurun_out/r2h_bench_reference.json; echo
ls -la gpurun_out | grep r2h
