#!/bin/bash
# tools/gpu_round.sh <tag> — one GPU session: GPU tests, smoke, bench (N = 1).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt; nproc >> gpurun_out/${tag}_gpu.txt; free -g >> gpurun_out/${tag}_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
