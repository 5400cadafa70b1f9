#!/bin/bash
# round 2, GPU call Y: sustained run (1500 steps = 10 s of timed region, clocks sampled), launch list of one step at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --no-extras --no-cpu --steps 1500 --warmup 3 > gpurun_out/r2y_bench_sustained.json 2> gpurun_out/r2y_bench_sustained.err; echo "sustained exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2y_bench_sustained.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["steps"], d["clocks"])
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2y_launches.csv $CMD > /dev/null 2>&1; echo "ncu exit $?"
