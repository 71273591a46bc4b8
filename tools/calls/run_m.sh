#!/bin/bash
# round 1, call m: refit / quality / image-file tests + dynamic-scene bench + one bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r01m_pytest_gpu.log
tail -15 $OUT/r01m_pytest_gpu.log
timeout 600 python tools/bench_dynamic.py --reps 5 > $OUT/r01m_dynamic.json 2> $OUT/r01m_dynamic.err; tail -c 2500 $OUT/r01m_dynamic.json; tail -5 $OUT/r01m_dynamic.err
timeout 600 python bench.py --steps 6 --no-cpu-baseline > $OUT/r01m_bench.json 2> $OUT/r01m_bench.err; tail -c 1500 $OUT/r01m_bench.json
