#!/bin/bash
# One gpurun call = tests + bench (both arms) + ncu launch list + ncu full capture of the traversal kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests] ; everything lands in gpurun_out/<tag>_*
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
# full capture of one closest-hit and one occlusion launch on the configs[1] streams
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trace -c 2 \
  -f -o $OUT/${TAG}_trace python tools/profile_trace.py --workload c2 --reps 1 > $OUT/${TAG}_ncu_trace.log 2>&1
tail -3 $OUT/${TAG}_ncu_trace.log
# configs[2]-sized scene (10 M triangles): timing only
timeout 900 python tools/profile_trace.py --workload c3 --reps 3 > $OUT/${TAG}_c3.log 2>&1
tail -5 $OUT/${TAG}_c3.log
