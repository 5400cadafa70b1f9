#!/bin/bash
# round 2, GPU call F2: SM clocks sampled during the end-to-end loop as well
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu --steps 150 --warmup 3"
timeout 300 $B > gpurun_out/r2f2_bench.json 2>/dev/null; echo "exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f2_bench.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["clocks"], d["e2e"]["clocks"])
PY
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm --format=csv
