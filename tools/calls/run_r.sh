#!/bin/bash
# occupancy A/B of the traversal kernels (launch bounds 8 / 9 / 10 CTAs per SM) + shard balance of the bench batch at world 8
OUT=gpurun_out; mkdir -p $OUT
{
for W in c3 c2; do
  for R in 1 2; do
    echo "== $W default (8 CTAs/SM, run $R)"; python tools/profile_trace.py --workload $W --reps 3 | grep closest
    for V in ctas9 ctas10; do
      echo "== $W $V (run $R)"; python tools/profile_trace.py --workload $W --reps 3 --lib embree-aarch64_b200/lib/variants/libembree3_$V.so | grep closest
    done
  done
done
} > $OUT/r02l_ab_ctas.log 2>&1
python tools/check_variant.py embree-aarch64_b200/lib/variants/libembree3_ctas9.so >> $OUT/r02l_ab_ctas.log 2>&1
python tools/shard_balance.py --world 8 > $OUT/r02l_shard_balance.jsonl 2> $OUT/r02l_shard_balance.err
tail -5 $OUT/r02l_shard_balance.jsonl
