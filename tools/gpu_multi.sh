#!/bin/bash
# round 2, GPU call H2 (N GPUs, HEAD of round 2): NCCL exchange steps through the C ABI against the single-GPU results on 4 ranks, bench under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi topo -m > gpurun_out/r2h2_topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2h2_multi_gpu_check_${N}gpu.log 2>&1; echo "multi_gpu_check exit $?"
tail -4 gpurun_out/r2h2_multi_gpu_check_${N}gpu.log
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2h2_bench_${N}gpu.json 2> gpurun_out/r2h2_bench_${N}gpu.err; echo "bench ${N}gpu exit $?"
tail -c 600 gpurun_out/r2h2_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h2_bench_${N}gpu.json"))
    print(round(d["value"]), round(d["e2e"]["value"]), d["n_gpus"])
    for k,v in (d.get("configs") or {}).items(): print(k, json.dumps(v)[:700])
except Exception as e: print("failed", e)
PY
