#!/bin/bash
# Multi-GPU pass (N = $1): gpu tests, in-process sweep, torchrun bench (both arms).
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu_n$N.log 2>&1
echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu_n$N.log
rm -f gpurun_out/sweep_n$N.jsonl
timeout 600 python tools/sweep.py --gpus $N --ctas ${CTAS:-148,111,74,48,32,16} --iters 5 --overlap 0,1 --uni 0,1 --paths ${PATHS:-0,1} --out gpurun_out/sweep_n$N.jsonl > gpurun_out/sweep_n$N.log 2>&1
echo "sweep exit=$?"; tail -2 gpurun_out/sweep_n$N.log
PORT=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --impl reference --gpus $N --steps 20 --warmup 2 > gpurun_out/ref_n$N.json 2> gpurun_out/ref_n$N.err
echo "ref exit=$?"; cat gpurun_out/ref_n$N.json
