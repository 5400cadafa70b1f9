#!/bin/bash
# round 2, GPU call G (8 GPUs): the NCCL exchange steps and the bench under torchrun at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
nproc > gpurun_out/r2g_nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tests/multi_gpu_check.py > gpurun_out/r2g_multi_gpu_check.out 2>&1; echo "multi_gpu_check exit $?"
tail -2 gpurun_out/r2g_multi_gpu_check.out | cut -c1-600
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2g_bench_8gpu.json 2> gpurun_out/r2g_bench_8gpu.err; echo "bench 8gpu exit $?"
tail -c 600 gpurun_out/r2g_bench_8gpu.err
ls -la gpurun_out | grep r2g
