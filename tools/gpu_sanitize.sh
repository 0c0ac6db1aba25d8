#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_DONE" gpurun_out/sanitizer_$tool.log | head -4
done
