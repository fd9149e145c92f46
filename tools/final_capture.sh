#!/bin/bash
# tools/final_capture.sh <tag> — one single-GPU session that produces what profiles/ cites for a round: GPU tests, smoke, the bench
# (both arms), the ncu launch list of the bench command, ncu captures of the C3 kernels (whole frame, and one rank's share of 8),
# and the parity report.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -c 400 $O/bench_$TAG.json; echo
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "reference arm rc=$?"; tail -c 300 $O/bench_ref_$TAG.json; echo
# the launch list of the bench command itself (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > $O/ncu_bench.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
run() {   # name, world, opts, launches per frame
  export C3_WORLD=$2 C3_PHASE=0 S3R_OPTS=$3
  timeout 600 ncu --metrics $M --clock-control none --launch-skip $(( $4 * 15 )) -c $4 --csv --log-file $O/${TAG}_l_$1.csv python tools/c3_band_probe.py > /dev/null 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none --launch-skip $(( $4 * 9 )) -c $4 -f -o $O/${TAG}_c3_$1 python tools/c3_band_probe.py > /dev/null 2>&1
}
run w1raw 1 clusters=0 6
run w8 8 clusters=1 7
unset S3R_OPTS C3_WORLD C3_PHASE
timeout 600 python tests/parity_report.py --out $O/parity_$TAG 2>&1 | tail -5
ls -la $O/${TAG}_c3_*.ncu-rep $O/launches_bench_$TAG.csv
