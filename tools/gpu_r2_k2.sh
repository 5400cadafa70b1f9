#!/bin/bash
# round 2, GPU call K2: the GPU suite three times over (flakiness check of the flag-based integral kernel, the warp-specialised matcher and the graph path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -p no:randomly > gpurun_out/r2k2_pytest_$r.log 2>&1; echo "run $r exit $?"
  tail -1 gpurun_out/r2k2_pytest_$r.log
done
