#!/bin/bash
# tools/gpu_sortmiddle_ab.sh <tag> — C3 whole frame on one GPU through the routings of the general path:
# visibility keys + record-free direct walk (default) | a setup record per candidate + flat walk (direct_small=0) |
# the same with everything from 16 x 16 pixels up through the sort-middle tile bins (flat_max=16)
tag=${1:-sm}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "general_path_variants" 2>&1 | tail -3
for opts in "clusters=0" "clusters=0,direct_small=0" "clusters=0,direct_small=0,flat_max=16"; do
  echo "== $opts" | tee -a gpurun_out/${tag}_sortmiddle.log
  C3_WORLD=1 C3_PHASE=0 C3_WARM=8 S3R_OPTS=$opts timeout 600 python tools/c3_band_probe.py 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/${tag}_sortmiddle.log
done
