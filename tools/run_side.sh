#!/bin/bash
# HEAD: the side benches behind DESIGN 4.6 (configs[0], dynamic scenes, instancing), each with the reference beside it
OUT=gpurun_out; TAG=${1:-r02final}; mkdir -p $OUT
timeout 400 python tools/bench_c1.py > $OUT/${TAG}_c1.json 2> $OUT/${TAG}_c1.err; tail -c 400 $OUT/${TAG}_c1.json
timeout 400 python tools/bench_dynamic.py > $OUT/${TAG}_dynamic.json 2> $OUT/${TAG}_dynamic.err; tail -c 400 $OUT/${TAG}_dynamic.json
timeout 400 python tools/bench_instancing.py > $OUT/${TAG}_instancing.json 2> $OUT/${TAG}_instancing.err; tail -c 400 $OUT/${TAG}_instancing.json
