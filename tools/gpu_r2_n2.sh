#!/bin/bash
# round 2, GPU call N2: GPU suite + smoke after the oracle's switches (the checker's defaults are unchanged)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/r2n2_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/r2n2_pytest_gpu.log
