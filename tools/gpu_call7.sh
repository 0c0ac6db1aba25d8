#!/bin/bash
# N-GPU "final profile" pass: gpu tests, bench (both arms), sweeps (paths 0/1/2, uni, full), trace, storm.
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu_n${N}_c7.log 2>&1
echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu_n${N}_c7.log
PORT=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n${N}_c7.json 2> gpurun_out/bench_n${N}_c7.err
echo "bench exit=$?"; cat gpurun_out/bench_n${N}_c7.json | cut -c1-3000; tail -3 gpurun_out/bench_n${N}_c7.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --impl reference --gpus $N --steps 10 --warmup 1 > gpurun_out/ref_n${N}_c7.json 2> gpurun_out/ref_n${N}_c7.err
echo "ref exit=$?"; cat gpurun_out/ref_n${N}_c7.json | cut -c1-600
rm -f gpurun_out/sweep_n${N}_c7.jsonl gpurun_out/sweep_full_n${N}_c7.jsonl
timeout 600 python tools/sweep.py --gpus $N --ctas 148,111 --iters 5 --overlap 1 --uni 0,1 --paths 0,2 --out gpurun_out/sweep_n${N}_c7.jsonl > gpurun_out/sweep_n${N}_c7.log 2>&1
echo "sweep exit=$?"; tail -2 gpurun_out/sweep_n${N}_c7.log | cut -c1-400
timeout 600 python tools/sweep.py --gpus $N --mode full --ctas 148 --iters 3 --overlap 1 --uni 0 --paths 0 --out gpurun_out/sweep_full_n${N}_c7.jsonl > gpurun_out/sweep_full_n${N}_c7.log 2>&1
echo "full sweep exit=$?"; cat gpurun_out/sweep_full_n${N}_c7.jsonl | cut -c1-600
timeout 300 python tools/trace.py --gpus $N --out gpurun_out/trace_n${N}_c7.json > gpurun_out/trace_n${N}_c7.txt 2>&1; echo "trace exit=$?"; head -20 gpurun_out/trace_n${N}_c7.txt
timeout 600 python tools/storm.py --gpus $N --cycles ${STORM:-200} > gpurun_out/storm_n${N}_c7.json 2> gpurun_out/storm_n${N}_c7.err; echo "storm exit=$?"; cat gpurun_out/storm_n${N}_c7.json
