#!/bin/bash
# development aid: GPU test-suite + per-kernel time/instruction list of the C3 frame (whole frame and the 1/8 interleaved share)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in 1 8; do
  export C3_WORLD=$w C3_PHASE=0
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum --clock-control none --launch-skip 72 -c 6 --csv --log-file gpurun_out/l_w$w.csv python tools/c3_band_probe.py > /dev/null 2>&1
done
