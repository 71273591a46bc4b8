#!/bin/bash
# Sweeps traversal schedule knobs (device config string) and experiment builds on the configs[1] streams; timing only.
# usage: tools/sweep_trace.sh <tag> "<cfg1>" "<cfg2>" ...   (a cfg starting with lib= selects embree-aarch64_b200/lib/variants/libembree3_<name>.so)
OUT=gpurun_out; TAG=${1:-sweep}; shift; mkdir -p $OUT
for cfg in "$@"; do
  lib=""; c="$cfg"
  if [[ "$cfg" == lib=* ]]; then name="${cfg#lib=}"; name="${name%%,*}"; lib="--lib embree-aarch64_b200/lib/variants/libembree3_${name}.so"; c="${cfg#lib=$name}"; c="${c#,}"; fi
  echo "== $cfg" >> $OUT/${TAG}_sweep.log
  timeout 300 python tools/profile_trace.py --workload ${WORKLOAD:-c2} --reps 3 $lib --cfg "$c" 2>&1 | tail -1 >> $OUT/${TAG}_sweep.log
done
cat $OUT/${TAG}_sweep.log
