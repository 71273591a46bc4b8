#!/bin/bash
# emission rewritten (8 lanes per node, device-driven levels), treelet DP as its own level-synchronous kernel: tests, memcheck, build phases, ncu
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02n_pytest_gpu.log; tail -4 $OUT/r02n_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > $OUT/r02n_memcheck.log 2>&1; tail -3 $OUT/r02n_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > $OUT/r02n_racecheck.log 2>&1; tail -3 $OUT/r02n_racecheck.log
python tools/profile_build.py --workload c3 > $OUT/r02n_build_c3.jsonl 2>&1; cat $OUT/r02n_build_c3.jsonl
python tools/profile_build.py --workload c2 > $OUT/r02n_build_c2.jsonl 2>&1
python tools/profile_build.py --workload c3 --cfg gpu_builder=lbvh > $OUT/r02n_build_c3_lbvh.jsonl 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit$|k_treelet_dp|k_treelet_build" -c 16 -f -o $OUT/r02n_build \
  python tools/profile_build.py --workload c3 --commits 1 > $OUT/r02n_ncu_build.log 2>&1
tail -2 $OUT/r02n_ncu_build.log
