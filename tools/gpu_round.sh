#!/bin/bash
# tools/gpu_round.sh <tag> [bench args] — one GPU session: probes, GPU tests, smoke, bench (N = 1).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt; nproc >> gpurun_out/${tag}_gpu.txt; free -g >> gpurun_out/${tag}_gpu.txt
if [ -x tools/probes/red_probe ]; then timeout 120 tools/probes/red_probe > gpurun_out/${tag}_red_probe.txt 2>&1; cat gpurun_out/${tag}_red_probe.txt; fi
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 "$@" ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"] and d["e2e"]["value"], "banded", d["banded_equals_whole"], "cpu==b200", d["cpu_baseline"] and d["cpu_baseline"]["frame0_equals_b200"])
    for k, v in d["roofline"]["kernels"].items(): print("  %-20s %8.1f us" % (k, v["avg_launch_us"]))
    print("secondary", d["secondary"] and (d["secondary"]["value"], d["secondary"]["e2e"] and d["secondary"]["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/${tag}_bench.err
