#!/bin/bash
# tools/gpu_session.sh <tag> — GPU tests, C3 probes (whole frame and a 1/8 share), bench N = 1 (no CPU legs), ncu launch lists.
tag=${1:-r02}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for w in 1 8; do C3_WORLD=$w C3_PHASE=0 timeout 300 python tools/c3_band_probe.py 2>&1 | tail -1; done
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary > $O/${tag}_bench.json 2> $O/${tag}_bench.err; tail -3 $O/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("$O/${tag}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["banded_equals_whole"])
for k, v in d["roofline"]["kernels"].items(): print("  %-20s %8.1f us" % (k, v["avg_launch_us"]))
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
for w in 1 8; do
  C3_WORLD=$w C3_PHASE=0 timeout 600 ncu --metrics $M --clock-control none --launch-skip 84 -c 7 --csv --log-file $O/${tag}_l_w$w.csv python tools/c3_band_probe.py > /dev/null 2>&1
  echo "== w$w"; grep "gpu__time_duration\|inst_executed.sum\|issue_active" $O/${tag}_l_w$w.csv | awk -F'","' '{printf "%-30s %-52s %s\n", substr($5,1,28), $13, $NF}'
done
