#!/bin/bash
# 1-GPU measurement pass: sweep, bench (both arms), ncu launch list, one ncu --set full capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py --gpus 1 --ctas 148,111,74,48 --iters 6 > gpurun_out/sweep1.log 2>&1
echo "sweep exit=$?"; tail -3 gpurun_out/sweep1.log
timeout 600 python bench.py --gpus 1 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench exit=$?"; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
timeout 300 python bench.py --gpus 1 --path ldst --no-cpu-baseline > gpurun_out/bench1_ldst.json 2>> gpurun_out/bench1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 2 > gpurun_out/ref1.json 2> gpurun_out/ref1.err
echo "ref exit=$?"; cat gpurun_out/ref1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n1.csv \
    python bench.py --gpus 1 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit=$?"; tail -3 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdprobe_kernel -s 4 -c 1 -f -o gpurun_out/prof_n1 \
    python bench.py --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
