#!/bin/bash
# Final per-N evidence: gpu tests + both bench arms (+ trace). N=1 additionally: ncu launch list and one --set full capture.
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider > gpurun_out/final_pytest_n$N.log 2>&1
echo "pytest exit=$?"; tail -3 gpurun_out/final_pytest_n$N.log
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench exit=$?"
  timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 2 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err; echo "ref exit=$?"
  timeout 300 python bench.py --gpus 1 --path ldst --no-cpu-baseline > gpurun_out/final_bench_n1_ldst.json 2>> gpurun_out/final_bench_n1.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_n1.csv \
      python bench.py --gpus 1 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_launch.log 2>&1; echo "ncu launches exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdprobe_kernel -s 4 -c 1 -f -o gpurun_out/final_prof_n1 \
      python bench.py --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full exit=$?"
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke exit=$?"; cat gpurun_out/final_smoke.log
else
  PORT=$((20000 + RANDOM % 20000))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench exit=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
      bench.py --impl reference --gpus $N --steps 10 --warmup 1 > gpurun_out/final_ref_n$N.json 2> gpurun_out/final_ref_n$N.err; echo "ref exit=$?"
  timeout 300 python tools/trace.py --gpus $N --out gpurun_out/final_trace_n$N.json > gpurun_out/final_trace_n$N.txt 2>&1; echo "trace exit=$?"
fi
cat gpurun_out/final_bench_n$N.json | cut -c1-1500
