#!/bin/bash
# round 2, GPU call L2: exact kernel for train sets under 128 rows in the generic matcher call: full GPU suite, smoke, short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l2_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2l2_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2l2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l2_pytest_gpu.log
tail -4 gpurun_out/r2l2_pytest_gpu.log
timeout 300 python bench.py --no-extras --no-cpu --steps 30 --warmup 3 > gpurun_out/r2l2_bench.json 2>/dev/null; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l2_bench.json"))
print(round(d["value"]), round(d["e2e"]["value"]))
PY
