#!/bin/bash
# NVLink byte counters of the solo probe kernel under ncu (2-GPU box, one process).  Output: gpurun_out/${TAG}_ncu_nvlink_{read,write}.csv
TAG=${TAG:-r02}
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
M=nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvltx__bytes_packet_request.sum,nvltx__bytes_packet_response.sum
M=$M,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvlrx__bytes_packet_request.sum,nvlrx__bytes_packet_response.sum
M=$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum
for leg in read:1 write:2; do
  name=${leg%%:*}; ops=${leg##*:}
  timeout 600 ncu --metrics $M --clock-control none -k regex:cdprobe_kernel --launch-skip 3 --launch-count 1 --csv \
      --log-file gpurun_out/${TAG}_ncu_nvlink_$name.csv python tools/solo_profile.py --ops $ops > gpurun_out/${TAG}_ncu_nvlink_$name.log 2>&1
  echo "ncu $name exit=$?"; tail -2 gpurun_out/${TAG}_ncu_nvlink_$name.log; grep -c nvl gpurun_out/${TAG}_ncu_nvlink_$name.csv
done
# the same two kernels unprofiled, for the GB/s that go with the byte counts
python tools/solo_profile.py --ops 1; python tools/solo_profile.py --ops 2
