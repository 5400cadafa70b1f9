#!/bin/bash
# round 2, GPU call O: software-pipelined key kernel (k_tc_candidates3p): matcher tests, A/B against the serial key kernel, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider -k "match or tensor or stream or db or correspond or knn or loop" > gpurun_out/r2o_pytest_match.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest_match.log
tail -5 gpurun_out/r2o_pytest_match.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
for r in 1 2; do
  timeout 300 $B > gpurun_out/r2o_bench_pipe$r.json 2> gpurun_out/r2o_bench_pipe$r.err; echo "pipe$r exit $?"
  SURF_TC_KEYS_SERIAL=1 timeout 300 $B > gpurun_out/r2o_bench_serial$r.json 2>/dev/null; echo "serial$r exit $?"
done
python - <<'PY'
import json
for n in ("pipe1","serial1","pipe2","serial2"):
    try:
        d=json.load(open(f"gpurun_out/r2o_bench_{n}.json"))
        st={s["stage"]:round(s["ms"],3) for s in d["roofline_stages"]}
        print(n, round(d["value"]), round(d["e2e"]["value"]), st)
    except Exception as e: print(n, "failed", e)
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_' -s 12 -c 6 -o gpurun_out/r2o_match $CMD > gpurun_out/r2o_ncu_match.log 2>&1; echo "ncu exit $?"
