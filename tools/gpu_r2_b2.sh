#!/bin/bash
# round 2, GPU call B2: k_orient sample constants in registers; redo grouping as before
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider -x > gpurun_out/r2b2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b2_pytest_gpu.log
tail -6 gpurun_out/r2b2_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
for r in 1 2; do
  timeout 300 $B > gpurun_out/r2b2_bench$r.json 2> gpurun_out/r2b2_bench$r.err; echo "bench$r exit $?"
done
python - <<'PY'
import json
for n in ("1","2"):
    try:
        d=json.load(open(f"gpurun_out/r2b2_bench{n}.json"))
        st={s["stage"]:round(s["ms"],4) for s in d["roofline_stages"]}
        print(n, round(d["value"]), round(d["e2e"]["value"]), st)
    except Exception as e: print(n, "failed", e)
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_tc_redo' -s 4 -c 4 -o gpurun_out/r2b2_orient $CMD > gpurun_out/r2b2_ncu.log 2>&1; echo "ncu exit $?"
