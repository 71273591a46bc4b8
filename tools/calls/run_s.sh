#!/bin/bash
# builder changes (warp-aggregated emission counters, DP fused into the treelet kernel) + 9 CTAs/SM closest-hit: tests, build phases, ncu of the builder kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02m_pytest_gpu.log; tail -4 $OUT/r02m_pytest_gpu.log
python tools/profile_build.py --workload c3 > $OUT/r02m_build_c3.jsonl 2>&1; cat $OUT/r02m_build_c3.jsonl
python tools/profile_build.py --workload c2 > $OUT/r02m_build_c2.jsonl 2>&1
python tools/profile_trace.py --workload c3 --reps 3 | grep closest > $OUT/r02m_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_treelet_build|k_ploc_tail" -c 14 -f -o $OUT/r02m_build \
  python tools/profile_build.py --workload c3 --commits 1 > $OUT/r02m_ncu_build.log 2>&1
tail -2 $OUT/r02m_ncu_build.log
