#!/bin/bash
# round 2, GPU call I: parity + A/B (HEAD library vs the describe-stage changes) + source-level ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q --timeout=240 -rf -p no:cacheprovider > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest_gpu.log
tail -4 gpurun_out/r2i_pytest_gpu.log
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
for r in 1 2; do
  timeout 300 $B > gpurun_out/r2i_bench_new$r.json 2> gpurun_out/r2i_bench_new$r.err; echo "new$r exit $?"
  SURF_B200_LIB=$PWD/scratch_ab/lib_head.so timeout 300 $B --dev-lib > gpurun_out/r2i_bench_head$r.json 2>/dev/null; echo "head$r exit $?"
done
python - <<'PY'
import json
for n in ("new1","head1","new2","head2"):
    try:
        d=json.load(open(f"gpurun_out/r2i_bench_{n}.json"))
        st={s["stage"]:round(s["ms"],3) for s in d["roofline_stages"]}
        print(n, round(d["value"]), round(d["e2e"]["value"]), st)
    except Exception as e: print(n, "failed", e)
PY
CMD="python bench.py --no-extras --no-cpu --steps 2 --warmup 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_orient|k_descriptor' -s 15 -c 5 -o gpurun_out/r2i_describe $CMD > gpurun_out/r2i_ncu_describe.log 2>&1; echo "ncu describe exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hessian_nms_tma_c|k_tc_candidates' -s 9 -c 3 -o gpurun_out/r2i_detect $CMD > gpurun_out/r2i_ncu_detect.log 2>&1; echo "ncu detect exit $?"
ls -la gpurun_out | grep r2i
