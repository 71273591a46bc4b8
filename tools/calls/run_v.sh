#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash tools/run_ab_trace.sh r02p embree-aarch64_b200/lib/variants/libembree3_fold.so > $OUT/r02p_ab_fold.log 2>&1
python tools/check_variant.py embree-aarch64_b200/lib/variants/libembree3_fold.so >> $OUT/r02p_ab_fold.log 2>&1
tail -30 $OUT/r02p_ab_fold.log
