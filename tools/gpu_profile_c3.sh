#!/bin/bash
# tools/gpu_profile_c3.sh <tag> — ncu captures of the C3 frame's kernels: whole frame on one GPU and one rank's share of 8.
tag=${1:-r02}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
for w in 1 8; do
  export C3_WORLD=$w C3_PHASE=0
  timeout 300 python tools/c3_band_probe.py > $O/${tag}_c3_w${w}_plain.log 2>&1; tail -1 $O/${tag}_c3_w${w}_plain.log
  # 3 warm-up rounds of 4 frames x 5 launches, then the timed 4 frames: skip into the last round
  timeout 600 ncu --metrics $M --clock-control none --launch-skip 60 -c 5 --csv --log-file $O/${tag}_l_w$w.csv python tools/c3_band_probe.py > /dev/null 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"cluster_front|shade_tiles" --launch-skip 24 -c 2 -f -o $O/${tag}_c3_w$w python tools/c3_band_probe.py > /dev/null 2>&1
done
ls -la $O/${tag}_c3_w*.ncu-rep
