#!/bin/bash
# HEAD: compute-sanitizer (memcheck, racecheck), ncu of the builder kernels, two-bounce streams of configs[2]
OUT=gpurun_out; TAG=${1:-r02final}; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > $OUT/${TAG}_memcheck.log 2>&1; tail -2 $OUT/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > $OUT/${TAG}_racecheck.log 2>&1; tail -2 $OUT/${TAG}_racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit$|k_treelet_dp|k_treelet_build|k_ploc_tail" -c 16 -f -o $OUT/${TAG}_build \
  python tools/profile_build.py --workload c3 --commits 1 > $OUT/${TAG}_ncu_build.log 2>&1
timeout 900 python tools/bench_bounces.py > $OUT/${TAG}_c3_bounces.json 2> $OUT/${TAG}_bounces.err; tail -c 600 $OUT/${TAG}_c3_bounces.json
