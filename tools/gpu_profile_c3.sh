#!/bin/bash
# tools/gpu_profile_c3.sh <tag> — ncu captures of the C3 frame's kernels: whole frame on one GPU (raw-stream front and cluster
# front) and one rank's share of 8 (interleaved tile rows; cluster front).  Launch lists (gpu__time_duration + a few counters)
# and one --set full capture per shape.
tag=${1:-r02}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
run() {   # name, world, opts, launches per frame
  export C3_WORLD=$2 C3_PHASE=0 S3R_OPTS=$3
  timeout 300 python tools/c3_band_probe.py > $O/${tag}_c3_$1_plain.log 2>&1; tail -1 $O/${tag}_c3_$1_plain.log | cut -c1-260
  # 3 warm-up rounds + the timed round of 4 frames: skip to the last frame of the last round
  timeout 600 ncu --metrics $M --clock-control none --launch-skip $(( $4 * 15 )) -c $4 --csv --log-file $O/${tag}_l_$1.csv python tools/c3_band_probe.py > /dev/null 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none --launch-skip $(( $4 * 9 )) -c $4 -f -o $O/${tag}_c3_$1 python tools/c3_band_probe.py > /dev/null 2>&1
}
run w1raw 1 clusters=0 6
run w1 1 clusters=1 7
run w8 8 clusters=1 7
ls -la $O/${tag}_c3_*.ncu-rep
