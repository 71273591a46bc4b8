#!/bin/bash
# One gpurun call = tests + bench (both arms) + ncu launch list + ncu full capture of the traversal kernels on the north-star scene.
# usage: tools/gpu_round.sh <tag> [skip-tests] ; everything lands in gpurun_out/<tag>_*  (then: python tools/summarize_profiles.py <tag>)
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline > $OUT/${TAG}_bench_c2.json 2>> $OUT/${TAG}_bench.err
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1
# full capture of one closest-hit and one occlusion launch over the 16.7 M-ray streams of the north-star scene (and of configs[1])
for W in c3 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trace -c 2 \
    -f -o $OUT/${TAG}_trace_$W python tools/profile_trace.py --workload $W --reps 1 --meta $OUT/${TAG}_trace_$W.meta.json > $OUT/${TAG}_ncu_trace_$W.log 2>&1
  tail -3 $OUT/${TAG}_ncu_trace_$W.log
done
# shard balance of the bench batch at world 8 (one GPU plays every rank) and the build benchmark of configs[4] without the reference
timeout 600 python tools/shard_balance.py --world 8 --band 16 > $OUT/${TAG}_shard_balance.jsonl 2> $OUT/${TAG}_shard_balance.err
timeout 900 python bench.py --workload c5 --no-cpu-baseline --kinds scene > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
