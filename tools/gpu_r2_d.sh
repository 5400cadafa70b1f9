#!/bin/bash
# round 2, GPU call D (2 GPUs): NCCL exchange steps through the C ABI against the single-GPU results, bench under torchrun,
# then single-GPU probes (PCIe, upload policy)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2d_multi_gpu_check.out 2>&1; echo "multi_gpu_check exit $?"
tail -3 gpurun_out/r2d_multi_gpu_check.out
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo "bench 2gpu exit $?"
tail -c 800 gpurun_out/r2d_bench_2gpu.err
timeout 600 $TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d_bench_2gpu_reference.json 2> gpurun_out/r2d_bench_2gpu_reference.err; echo "reference arm exit $?"
python tools/pcie_probe.py > gpurun_out/r2d_pcie_probe.json 2>&1
B="python bench.py --no-extras --no-cpu --steps 30 --warmup 3"
SURF_B200_EARLY_UPLOAD=0 timeout 300 $B > gpurun_out/r2d_bench_upload_late.json 2>/dev/null; echo "late exit $?"
SURF_B200_EARLY_UPLOAD=1 timeout 300 $B > gpurun_out/r2d_bench_upload_early.json 2>/dev/null; echo "early exit $?"
ls -la gpurun_out | grep r2d
