#!/bin/bash
# tools/gpu_ab.sh <tag> <lib>... — per-kernel times of a C3 frame for several builds of the library (S3R_LIB); "default" = the shipped one
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  if [ "$lib" = default ]; then unset S3R_LIB; else export S3R_LIB=$PWD/$lib; fi
  while read -r cfg; do
    [ -z "$cfg" ] && continue
    set -- $cfg
    echo "== $lib: world $1 phase $2 $3" | tee -a gpurun_out/${tag}_ab.log
    C3_WORLD=$1 C3_PHASE=$2 S3R_OPTS=$3 timeout 300 python tools/c3_band_probe.py 2>&1 | tail -1 | cut -c1-250 | tee -a gpurun_out/${tag}_ab.log
  done <<< "${AB_CFGS:-$'1 0 clusters=0\n1 0 clusters=1\n8 3 clusters=1'}"
done
