#!/bin/bash
# 1-GPU: compute-sanitizer passes + storm (config 5) on one device.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_DONE|========= (Error|Race|Hazard)" gpurun_out/sanitizer_$tool.log | head -8
done
timeout 600 python tools/storm.py --gpus 1 --cycles 1000 > gpurun_out/storm_n1.json 2> gpurun_out/storm_n1.err; echo "storm n1 exit=$?"; cat gpurun_out/storm_n1.json
timeout 600 python tools/storm.py --gpus 4 --same-device --bytes $((64<<20)) --cycles 300 > gpurun_out/storm_same4.json 2> gpurun_out/storm_same4.err; echo "storm same4 exit=$?"; cat gpurun_out/storm_same4.json; tail -3 gpurun_out/storm_same4.err
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu_c5.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu_c5.log
