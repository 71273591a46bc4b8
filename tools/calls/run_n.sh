#!/bin/bash
# round 1, call n: instancing tests + instancing bench
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_instancing.py -m gpu -x -q > $OUT/r01n_pytest_inst.log 2>&1; echo "pytest exit $?" >> $OUT/r01n_pytest_inst.log
tail -40 $OUT/r01n_pytest_inst.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r01n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r01n_pytest_gpu.log
tail -5 $OUT/r01n_pytest_gpu.log
timeout 600 python tools/bench_instancing.py > $OUT/r01n_instancing.json 2> $OUT/r01n_instancing.err; tail -c 2500 $OUT/r01n_instancing.json; tail -5 $OUT/r01n_instancing.err
