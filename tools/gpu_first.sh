#!/bin/bash
# First GPU shake-out: smoke, then the gpu-marked parity tests. Everything under `timeout`.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
ls /dev/nvidia-caps-imex-channels >> gpurun_out/gpus.txt 2>&1
nproc >> gpurun_out/gpus.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"
tail -60 gpurun_out/pytest_gpu.log
