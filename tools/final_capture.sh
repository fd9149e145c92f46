#!/bin/bash
# One GPU session that produces everything profiles/ cites for a round: tests, bench (both arms), ncu launch
# lists and full captures of the dominant kernels of C2, C3 (whole frame and a 1/8 share) and C4, configs on 1 GPU.
TAG=${1:-r01}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 600 $O/bench_$TAG.json; echo
python bench.py --impl reference > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2_$TAG.csv python bench.py --steps 1 --warmup 3 --frames 48 --views-per-launch 8 --no-e2e --no-cpu > $O/ncu_c2.log 2>&1
# 48 frames = 6 launches of 8 poses per step; the 21st tile_raster launch is inside the timed step
ncu --set full --import-source on --clock-control none -k regex:tile_raster --launch-skip 20 -c 1 -f -o $O/${TAG}_c2_tile_raster python bench.py --steps 1 --warmup 3 --frames 48 --views-per-launch 8 --no-e2e --no-cpu > $O/ncu_c2_full.log 2>&1
for w in 1 8; do
  export C3_WORLD=$w C3_PHASE=0
  ncu --metrics $M --clock-control none --launch-skip 72 -c 6 --csv --log-file $O/l_w$w.csv python tools/c3_band_probe.py > /dev/null 2>&1
  ncu --set full --import-source on --clock-control none -k regex:"vertex_stage|triangle_classify|shade_tiles" --launch-skip 36 -c 3 -f -o $O/${TAG}_c3_w$w python tools/c3_band_probe.py > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --launch-skip 72 -c 6 --csv --log-file $O/l_c4.csv python tools/c4_probe.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"triangle_setup|post_setup|tile_raster_queue|shade_tiles" --launch-skip 48 -c 4 -f -o $O/${TAG}_c4 python tools/c4_probe.py > /dev/null 2>&1
python tests/run_configs.py --config c1,c2,c3,c4,c5 2>&1 | grep "^{" | cut -c1-260
python tests/parity_report.py --out $O/parity_$TAG 2>&1 | tail -7
ls -la $O/*.ncu-rep
