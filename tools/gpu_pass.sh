#!/bin/bash
# One GPU pass on an N-GPU box (N = $1): gpu tests, per-phase trace, barrier-mode comparison, bench (both arms),
# optional extra bench configs ($2..: c2 c3-full c5).  Everything lands in gpurun_out/ with the tag $TAG.
# SKIP_TESTS=1 skips pytest, BENCH_ONLY=1 also skips traces and the barrier sweep, TESTS_K narrows pytest (-k).
#   gpurun --gpus 2 --timeout 900 -- 'TAG=r02 tools/gpu_pass.sh 2 c2'
N=${1:-1}; shift
TAG=${TAG:-r02}
STEPS=${STEPS:-100}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out/${TAG}
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider ${TESTS_K:+-k "$TESTS_K"} > ${O}_pytest_n$N.log 2>&1
  echo "pytest exit=$?"; tail -12 ${O}_pytest_n$N.log
fi
if [ "$N" -gt 1 ] && [ -z "$BENCH_ONLY" ]; then
  timeout 300 python tools/trace.py --gpus $N --out ${O}_trace_n$N.json > ${O}_trace_n$N.txt 2>&1; echo "trace exit=$?"
  timeout 300 python tools/trace.py --gpus $N --flags 0x400 --out ${O}_trace_allrank_n$N.json > ${O}_trace_allrank_n$N.txt 2>&1
  timeout 300 python tools/trace.py --gpus $N --idle 1.0 --out ${O}_trace_cold_n$N.json > ${O}_trace_cold_n$N.txt 2>&1
  timeout 600 python tools/sweep.py --gpus $N --ctas 148 --iters 9 --overlap 1 --uni 0,1 --paths 0 --barriers 0,2,1 --out ${O}_barriers_n$N.jsonl > ${O}_barriers_n$N.log 2>&1
  echo "barrier sweep exit=$?"; cat ${O}_barriers_n$N.jsonl | cut -c1-400
fi
PORT=$((20000 + RANDOM % 20000))
run_bench() {  # $1 = output tag, rest = bench args
  local tag=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N "$@" > ${O}_${tag}_n$N.json 2> ${O}_${tag}_n$N.err
  else
    timeout 900 python bench.py --gpus 1 "$@" > ${O}_${tag}_n$N.json 2> ${O}_${tag}_n$N.err
  fi
  echo "bench $tag exit=$?"; cut -c1-1500 ${O}_${tag}_n$N.json; tail -3 ${O}_${tag}_n$N.err
  PORT=$((PORT + 1))
}
run_bench bench --steps $STEPS --warmup 5
run_bench ref --impl reference --steps 20 --warmup 2
for cfg in "$@"; do
  case $cfg in
    c2) [ "$N" = 2 ] && run_bench bench_c2 --config c2 --steps $STEPS --warmup 5 ;;
    c3-full) run_bench bench_c3full --config c3-full --steps 20 --warmup 3 ;;
    c5) run_bench bench_c5 --config c5 --steps ${CYCLES:-1000} --warmup 3 ;;
  esac
done
