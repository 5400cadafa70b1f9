#!/bin/bash
# round 2, GPU call N: per-problem redo lists; matcher tests, bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider -k "match or tensor or stream or db or correspond or knn or loop" > gpurun_out/r2n_pytest_match.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest_match.log
tail -5 gpurun_out/r2n_pytest_match.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
for r in 1 2; do
  timeout 300 $B > gpurun_out/r2n_bench_keys$r.json 2> gpurun_out/r2n_bench_keys$r.err; echo "keys$r exit $?"
done
python - <<'PY'
import json
for n in ("keys1","keys2"):
    try:
        d=json.load(open(f"gpurun_out/r2n_bench_{n}.json"))
        st={s["stage"]:round(s["ms"],3) for s in d["roofline_stages"]}
        print(n, round(d["value"]), round(d["e2e"]["value"]), st)
    except Exception as e: print(n, "failed", e)
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tc_' -c 60 --csv --log-file gpurun_out/r2n_launches_match.csv $CMD > /dev/null 2>&1; echo "ncu exit $?"
grep -c k_tc gpurun_out/r2n_launches_match.csv
