#!/bin/bash
# round 2, GPU call D2: validation at HEAD as the driver runs it: smoke, full GPU suite, default bench, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d2_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2d2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2d2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d2_pytest_gpu.log
tail -8 gpurun_out/r2d2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2d2_bench_default.json 2> gpurun_out/r2d2_bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference > gpurun_out/r2d2_bench_reference.json 2> gpurun_out/r2d2_bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d2_bench_default.json"))
st={s["stage"]:(round(s["ms"],4), round(s.get("frac") or 0,4)) for s in d["roofline_stages"]}
print(round(d["value"]), round(d["e2e"]["value"]), st, d["latency_ms_batch1"], d["clocks"], d["gpu_launches"])
r=json.load(open("gpurun_out/r2d2_bench_reference.json"))
print("reference", r["value"], r["cpu_baseline"]["cores"], r["config"]["workload"]==d["config"]["workload"])
PY
