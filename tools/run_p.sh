#!/bin/bash
# call p: compact hit download: correctness + e2e A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hit_download or pinned" > $OUT/r01p_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r01p_pytest.log
tail -25 $OUT/r01p_pytest.log
bash tools/bench_ab.sh r01p "-" "d2h=3" "d2h=3,chunk_rays=524288" "d2h=3,chunk_rays=2097152"
