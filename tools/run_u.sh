#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01u_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r01u_pytest_gpu.log; tail -4 $OUT/r01u_pytest_gpu.log
timeout 600 python bench.py --steps 6 --no-cpu-baseline > $OUT/r01u_bench.json 2> $OUT/r01u_bench.err; tail -c 1800 $OUT/r01u_bench.json
timeout 600 python tools/bench_build.py --kinds scene --sizes 1,10,50 --no-reference 2>&1 | cut -c1-900
