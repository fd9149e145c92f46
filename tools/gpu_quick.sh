#!/bin/bash
# tools/gpu_quick.sh <tag> — short GPU session: GPU tests, then per-kernel times of a C3 frame (whole frame with either front, one share of 8).
tag=${1:-q}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
for cfg in "1 0 clusters=0" "1 0 clusters=1" "8 3 clusters=1"; do
  set -- $cfg
  echo "== world $1 phase $2 $3" | tee -a gpurun_out/${tag}_probe.log
  C3_WORLD=$1 C3_PHASE=$2 S3R_OPTS=$3 timeout 300 python tools/c3_band_probe.py 2>&1 | tail -14 | tee -a gpurun_out/${tag}_probe.log
done
