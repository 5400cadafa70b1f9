#!/bin/bash
# round 2, GPU call W: CUDA-graph replay of small batches + vectorised carry of the integral kernel: full GPU suite, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest_gpu.log
tail -25 gpurun_out/r2w_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2w_bench_default.json 2> gpurun_out/r2w_bench_default.err; echo "bench exit $?"
tail -3 gpurun_out/r2w_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2w_bench_default.json"))
st={s["stage"]:round(s["ms"],4) for s in d["roofline_stages"]}
print(round(d["value"]), round(d["e2e"]["value"]), st)
for k,v in (d.get("latency") or {}).items(): print(k, {x:v[x] for x in ("ms_per_call_median","eager_ms_per_call_median","graph_replays","calls","launches_per_call")})
PY
