#!/bin/bash
# round 2, GPU call J2: the file-driven flow (rows f1-f3) at HEAD: sweeps from .mha files + correspondences, asserted against the oracle inside the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench_surfcmd.py > gpurun_out/r2j2_bench_surfcmd.json 2> gpurun_out/r2j2_bench_surfcmd.err; echo "surfcmd exit $?"
tail -c 1500 gpurun_out/r2j2_bench_surfcmd.json
tail -3 gpurun_out/r2j2_bench_surfcmd.err
