#!/bin/bash
# tools/gpu_trace2.sh <tag> — 2-GPU box: where the wall time of the in-process multi-GPU drop-in goes (S3R_MULTI_TRACE)
tag=${1:-t2}
mkdir -p gpurun_out
export S3R_MULTI_TRACE=1
for dev in 0 0,1; do
  echo "== devices $dev" | tee -a gpurun_out/${tag}_trace.log
  timeout 600 python tools/multi_e2e_sweep.py $dev c2,c3 4 2>&1 | tail -4 | cut -c1-900 | tee -a gpurun_out/${tag}_trace.log
done
