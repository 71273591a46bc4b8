#!/bin/bash
# A/B of device options through bench.py on one box: tools/bench_ab.sh <tag> "<cfg>" ...   ("-" = defaults)
OUT=gpurun_out; TAG=$1; shift; mkdir -p $OUT
for cfg in "$@"; do
  c="$cfg"; [ "$c" == "-" ] && c=""
  lib=""
  if [[ "$c" == lib=* ]]; then name="${c#lib=}"; name="${name%%,*}"; lib="embree-aarch64_b200/lib/variants/libembree3_${name}.so"; c="${c#lib=$name}"; c="${c#,}"; fi
  RQ_B200_LIB="$lib" RQ_B200_CFG="$c" timeout 600 python bench.py --steps 6 --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s value %7.1f closest %7.1f occluded %7.1f e2e %6.1f frac %.3f build %.2f ms' % ('$cfg', d['value'], d['closest_mrays_per_s'], d['occluded_mrays_per_s'], d['e2e']['value'], d['roofline']['frac'], d['build']['msTotal']))" | tee -a $OUT/${TAG}_ab.log
done
