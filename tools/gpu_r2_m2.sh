#!/bin/bash
# round 2, GPU call M2: the default bench line and the reference arm at the final HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2m2_bench_default.json 2> gpurun_out/r2m2_bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference > gpurun_out/r2m2_bench_reference.json 2> gpurun_out/r2m2_bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m2_bench_default.json"))
st={s["stage"]:(round(s["ms"],4), round(s.get("frac") or 0,4)) for s in d["roofline_stages"]}
print(round(d["value"]), round(d["e2e"]["value"]), st, d["latency_ms_batch1"], d["clocks"], d["e2e"].get("clocks"))
r=json.load(open("gpurun_out/r2m2_bench_reference.json"))
print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
