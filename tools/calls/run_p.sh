#!/bin/bash
# call p: compact hit download: correctness + e2e A/B
OUT=gpurun_out; mkdir -p $OUT
nproc > $OUT/r01p_nproc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hit_download or pinned" > $OUT/r01p_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r01p_pytest.log
tail -25 $OUT/r01p_pytest.log
bash tools/bench_ab.sh r01p2 "-" "d2h=3" "d2h=3,scatter_threads=4" "d2h=3,scatter_threads=12" "d2h=3,scatter_threads=16" "d2h=3,chunk_rays=524288"
