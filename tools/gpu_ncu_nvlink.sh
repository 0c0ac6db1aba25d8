#!/bin/bash
# NVLink byte counters of the solo probe kernel under ncu (2-GPU box, one process).
# Output: gpurun_out/${TAG}_ncu_nvlink_{read,write}.csv (+ .log).  Tries kernel replay first, then application
# replay (no device-memory save/restore: the kernel touches peer memory ncu may not be able to snapshot).
TAG=${TAG:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
M=nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvltx__bytes_packet_request.sum,nvltx__bytes_packet_response.sum
M=$M,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvlrx__bytes_packet_request.sum,nvlrx__bytes_packet_response.sum
M=$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
try() {  # $1 = name, $2 = ops, rest = extra ncu flags
  local name=$1 ops=$2; shift 2
  timeout 300 ncu --metrics $M --clock-control none -k regex:cdprobe_kernel --launch-skip 3 --launch-count 1 --csv "$@" \
      --log-file gpurun_out/${TAG}_ncu_nvlink_$name.csv python tools/solo_profile.py --ops $ops > gpurun_out/${TAG}_ncu_nvlink_$name.log 2>&1
  local rc=$?
  echo "ncu $name [$*] exit=$rc, nvl rows: $(grep -c nvl gpurun_out/${TAG}_ncu_nvlink_$name.csv 2>/dev/null)"
  return $rc
}
for leg in read:1 write:2; do
  name=${leg%%:*}; ops=${leg##*:}
  try $name $ops || try $name $ops --replay-mode application || try $name $ops --replay-mode application --cache-control none \
    || { echo "== sanity: time only"; timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cdprobe_kernel --launch-skip 3 --launch-count 1 python tools/solo_profile.py --ops $ops 2>&1 | tail -5; }
  head -c 600 gpurun_out/${TAG}_ncu_nvlink_$name.log
done
# the same two kernels unprofiled, for the GB/s that go with the byte counts
python tools/solo_profile.py --ops 1; python tools/solo_profile.py --ops 2
