#!/bin/bash
# L2 experiments on the traversal kernels: streaming (evict-first) ray loads / hit stores, persisting window over the top levels of the node array
OUT=gpurun_out; mkdir -p $OUT
V=embree-aarch64_b200/lib/variants/libembree3_cs.so
{
for W in c3 c2; do
  echo "== $W default"; python tools/profile_trace.py --workload $W --reps 3 | grep closest
  echo "== $W streaming rays"; python tools/profile_trace.py --workload $W --reps 3 --lib $V | grep closest
  for MB in 16 48; do
    echo "== $W persist ${MB} MB"; RQ_L2_PERSIST_MB=$MB python tools/profile_trace.py --workload $W --reps 3 | grep closest
    echo "== $W persist ${MB} MB + streaming rays"; RQ_L2_PERSIST_MB=$MB python tools/profile_trace.py --workload $W --reps 3 --lib $V | grep closest
  done
done
python tools/check_variant.py $V
} > $OUT/r02s_ab_l2.log 2>&1
tail -40 $OUT/r02s_ab_l2.log
