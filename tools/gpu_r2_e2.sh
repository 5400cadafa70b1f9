#!/bin/bash
# round 2, GPU call E2: what the end-to-end loop loses against the device-resident one (diagnosis variants of the e2e loop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu --steps 40 --warmup 3"
for v in full no_desc no_d2h no_h2d; do
  if [ $v = full ]; then timeout 300 $B > gpurun_out/r2e2_$v.json 2>/dev/null; else SURF_BENCH_E2E_DIAG=$v timeout 300 $B > gpurun_out/r2e2_$v.json 2>/dev/null; fi
  echo "$v exit $?"
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 $B > gpurun_out/r2e2_conn32.json 2>/dev/null; echo "conn32 exit $?"
python - <<'PY'
import json
for n in ("full","no_desc","no_d2h","no_h2d","conn32"):
    try:
        d=json.load(open(f"gpurun_out/r2e2_{n}.json"))
        print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), d["e2e"]["d2h_bytes_per_step"])
    except Exception as e: print(n, "failed", e)
PY
