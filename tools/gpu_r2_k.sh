#!/bin/bash
# round 2, GPU call K: re-entry check at HEAD (full GPU suite + the default bench line as the driver runs it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest_gpu.log
tail -4 gpurun_out/r2k_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k_bench_default.json"))
st={s["stage"]:round(s["ms"],3) for s in d["roofline_stages"]}
print(round(d["value"]), round(d["e2e"]["value"]), st, d["latency_ms_batch1"], d["clocks"])
PY
