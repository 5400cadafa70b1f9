#!/bin/bash
# round 2, GPU call C2: quad tournament in the key epilogue: full GPU suite, A/B against listing every key, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -rf -p no:cacheprovider > gpurun_out/r2c2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c2_pytest_gpu.log
tail -8 gpurun_out/r2c2_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
for r in 1 2; do
  timeout 300 $B > gpurun_out/r2c2_bench_quad$r.json 2> gpurun_out/r2c2_bench_quad$r.err; echo "quad$r exit $?"
  SURF_TC_KEYS_PAIRS=1 timeout 300 $B > gpurun_out/r2c2_bench_pairs$r.json 2>/dev/null; echo "pairs$r exit $?"
done
python - <<'PY'
import json
for n in ("quad1","pairs1","quad2","pairs2"):
    try:
        d=json.load(open(f"gpurun_out/r2c2_bench_{n}.json"))
        st={s["stage"]:round(s["ms"],4) for s in d["roofline_stages"]}
        print(n, round(d["value"]), round(d["e2e"]["value"]), st)
    except Exception as e: print(n, "failed", e)
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_' -s 12 -c 6 -o gpurun_out/r2c2_match $CMD > gpurun_out/r2c2_ncu.log 2>&1; echo "ncu exit $?"
