#!/bin/bash
# tools/gpu_stride.sh <tag> — bench (N = 1, no CPU arm, no e2e, no secondary) with per-kernel events on every frame and on every third
tag=${1:-st}
mkdir -p gpurun_out
for st in 1 3; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-secondary --timing-stride $st > gpurun_out/${tag}_bench_s$st.json 2> gpurun_out/${tag}_bench_s$st.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_s$st.json").read().strip().splitlines()[-1])
print("stride $st value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), {k: round(v["avg_launch_us"], 1) for k, v in d["roofline"]["kernels"].items()}, "shares", {k: round(v["share_of_step"], 3) for k, v in d["roofline"]["kernels"].items()})
PY
done
