#!/bin/bash
# cheaper slot assignment, gaps-only memset, level-synchronous thread phase of the treelet builder: tests, memcheck, build phases
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02o_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02o_pytest_gpu.log; tail -4 $OUT/r02o_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > $OUT/r02o_memcheck.log 2>&1; tail -3 $OUT/r02o_memcheck.log
python tools/profile_build.py --workload c3 > $OUT/r02o_build_c3.jsonl 2>&1; cat $OUT/r02o_build_c3.jsonl
python tools/profile_build.py --workload c2 > $OUT/r02o_build_c2.jsonl 2>&1
python tools/profile_build.py --workload c3 --cfg sweep_bottom=1,treelet=512 > $OUT/r02o_build_c3_high.jsonl 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit$|k_treelet_dp|k_treelet_build" -c 16 -f -o $OUT/r02o_build \
  python tools/profile_build.py --workload c3 --commits 1 > $OUT/r02o_ncu_build.log 2>&1
tail -2 $OUT/r02o_ncu_build.log
