#!/bin/bash
# tools/gpu_multi.sh <tag> <n> — N-GPU session: the multi-process peer test, then bench.py under torchrun at N ranks.
tag=${1:-r02}; n=${2:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_multigpu_peer.py tests/test_gpu_parity.py::test_multi_gpu_drop_in_matches_the_single_gpu_frames -x -q 2>&1 | tail -5
for k in $(echo $n | tr ',' ' '); do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k bench.py --gpus $k --steps 10 --warmup 3 --no-cpu > $O/${tag}_bench_n$k.json 2> $O/${tag}_bench_n$k.err
  echo "bench N=$k rc=$?"; tail -4 $O/${tag}_bench_n$k.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("$O/${tag}_bench_n$k.json").read().strip().splitlines()[-1])
    print("N=$k value", round(d["value"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "gpus", d["e2e"] and d["e2e"]["gpus"], "banded_equals_whole", d["banded_equals_whole"])
    for kk, v in d["roofline"]["kernels"].items(): print("  %-20s %8.1f us" % (kk, v["avg_launch_us"]))
    s = d["secondary"]; print("  secondary", s and round(s["value"], 1), s and s["e2e"] and round(s["e2e"]["value"], 1))
except Exception as e:
    print("parse failed", e)
PY
done
