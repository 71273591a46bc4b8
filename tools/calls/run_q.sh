#!/bin/bash
# call q: host-packed upload + compact download: correctness + e2e A/B
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hit_download or pinned or full_size" > $OUT/r01q_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/r01q_pytest.log
tail -25 $OUT/r01q_pytest.log
bash tools/bench_ab.sh r01q "-" "pack_rays=0" "host_threads=8" "host_threads=12" "chunk_rays=524288" "chunk_rays=2097152" "d2h=0"
