#!/bin/bash
# The 1-GPU pass: what the driver runs at round end (pytest -m gpu, smoke, bench both arms) plus the ncu evidence
# for the N = 1 kernel: launch list (share of the step) and one `--set full` capture (dram traffic, source page).
#   gpurun --timeout 1500 -- 'TAG=r02 tools/gpu_n1.sh'            (SANITIZE=1 adds compute-sanitizer, slow)
TAG=${TAG:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
O=gpurun_out/${TAG}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 400 -p no:cacheprovider > ${O}_pytest_n1.log 2>&1
echo "pytest exit=$?"; tail -6 ${O}_pytest_n1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench exit=$?"; cut -c1-900 ${O}_bench_n1.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 2 > ${O}_ref_n1.json 2> ${O}_ref_n1.err; echo "ref exit=$?"
timeout 300 python tools/sweep.py --gpus 1 --ctas 148 --iters 9 --overlap 1 --paths 0,1,2 --out ${O}_sweep_n1.jsonl > ${O}_sweep_n1.log 2>&1; cut -c1-500 ${O}_sweep_n1.jsonl
# every launch of `bench.py --steps 3` with its device time: the probe is one kernel per step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file ${O}_launches_n1.csv \
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-daemon > ${O}_ncu_launch.log 2>&1; echo "ncu launch list exit=$?"
# one full capture of a steady-state launch (skip the open-time checksum kernel and the warm-up steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdprobe_kernel --launch-skip 6 --launch-count 1 \
    -o ${O}_prof_n1 -f python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-daemon > ${O}_ncu_full.log 2>&1; echo "ncu full exit=$?"
ncu -i ${O}_prof_n1.ncu-rep --page details --csv > ${O}_n1_ncu_details.csv 2>/dev/null
grep -E "dram__bytes_(read|write).sum|DRAM Throughput|Duration|Memory Throughput" ${O}_n1_ncu_details.csv | head -12
ncu -i ${O}_prof_n1.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum 2>/dev/null | tail -3 > ${O}_n1_ncu_raw.csv; cat ${O}_n1_ncu_raw.csv | cut -c1-600
if [ -n "$SANITIZE" ]; then
  for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py > ${O}_sanitizer_$tool.txt 2>&1
    echo "sanitizer $tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_DONE" ${O}_sanitizer_$tool.txt | head -4
  done
fi
