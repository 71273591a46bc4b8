#!/bin/bash
# same-box A/B of two builds of the library on the traversal streams: tools/run_ab_trace.sh <tag> <variant .so> [workloads]
TAG=$1; VAR=$2; shift 2
for W in ${@:-c3 c2}; do
  for R in 1 2; do
    echo "== $W default (run $R)"; python tools/profile_trace.py --workload $W --reps 3 | grep closest
    echo "== $W variant $VAR (run $R)"; python tools/profile_trace.py --workload $W --reps 3 --lib $VAR | grep closest
  done
done
