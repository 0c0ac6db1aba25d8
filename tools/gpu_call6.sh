#!/bin/bash
# N-GPU follow-up: bench (new defaults), traces, storm with real remap, config-2 and full-mode sweeps, daemon test.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PORT=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n${N}_b.json 2> gpurun_out/bench_n${N}_b.err
echo "bench exit=$?"; cat gpurun_out/bench_n${N}_b.json | cut -c1-2500; tail -3 gpurun_out/bench_n${N}_b.err
timeout 300 python tools/trace.py --gpus $N --out gpurun_out/trace_n$N.json > gpurun_out/trace_n$N.txt 2>&1; echo "trace exit=$?"; head -40 gpurun_out/trace_n$N.txt
timeout 300 python tools/trace.py --gpus $N --flags 0x100 --out gpurun_out/trace_n${N}_serial.json > gpurun_out/trace_n${N}_serial.txt 2>&1
timeout 600 python tools/storm.py --gpus $N --cycles 1000 > gpurun_out/storm_n$N.json 2> gpurun_out/storm_n$N.err; echo "storm exit=$?"; cat gpurun_out/storm_n$N.json; tail -2 gpurun_out/storm_n$N.err
rm -f gpurun_out/sweep_full_n$N.jsonl gpurun_out/sweep_cfg2.jsonl
timeout 600 python tools/sweep.py --gpus $N --mode full --ctas 148 --iters 5 --overlap 1 --uni 0,1 --paths 0 --out gpurun_out/sweep_full_n$N.jsonl > gpurun_out/sweep_full_n$N.log 2>&1; echo "full sweep exit=$?"; cat gpurun_out/sweep_full_n$N.jsonl | cut -c1-700
timeout 300 python tools/sweep.py --gpus 2 --mode full --bytes $((64<<20)) --ctas 148,74 --iters 8 --overlap 1 --uni 0,1 --paths 0,1 --out gpurun_out/sweep_cfg2.jsonl > gpurun_out/sweep_cfg2.log 2>&1; echo "cfg2 sweep exit=$?"; cat gpurun_out/sweep_cfg2.jsonl | cut -c1-600
timeout 300 python tools/sweep.py --gpus $N --mode reach --ctas 148,16 --iters 20 --overlap 1 --uni 0 --paths 0 --out gpurun_out/sweep_reach_n$N.jsonl > gpurun_out/sweep_reach_n$N.log 2>&1; cat gpurun_out/sweep_reach_n$N.jsonl | cut -c1-600
timeout 600 python -m pytest tests/test_daemon.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu_c6.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu_c6.log
