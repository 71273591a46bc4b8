#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02y_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02y_pytest_gpu.log; tail -4 $OUT/r02y_pytest_gpu.log
python tools/profile_build.py --workload c3 > $OUT/r02y_build_c3.jsonl 2>&1; tail -2 $OUT/r02y_build_c3.jsonl
python tools/profile_build.py --workload c3 --cfg gpu_builder=ploc > $OUT/r02y_build_c3_ploc.jsonl 2>&1; tail -2 $OUT/r02y_build_c3_ploc.jsonl
