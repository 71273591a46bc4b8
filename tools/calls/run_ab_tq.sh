#!/bin/bash
V=embree-aarch64_b200/lib/variants/libembree3_tq.so
python tools/check_variant.py $V
for W in c3 c2; do
  echo "== $W default"; python tools/profile_trace.py --workload $W --reps 3 --counters | grep "closest"
  for K in 12 16 20 24 28; do
    echo "== $W tq tvote=$K"; python tools/profile_trace.py --workload $W --reps 3 --lib $V --cfg tvote=$K --counters | grep "closest"
  done
done
