#!/bin/bash
# 8-GPU box, end-to-end (host streams) variants where the host memory system is the bound: default (compact hit list + host scatter),
# zerocopy=1 (kernel writes hits into the caller's page-locked records over PCIe), d2h=0 (whole records back by DMA), streaming-store scatter
OUT=gpurun_out; TAG=${1:-r02u}; mkdir -p $OUT
run() {  # name, lib, cfg
  RQ_B200_LIB="$2" RQ_B200_CFG="$3" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 3 --warmup 3 2> $OUT/${TAG}_$1.err | tail -1 > $OUT/${TAG}_$1.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_$1.json").read())
    print("$1", "value %.0f e2e %.1f pageable %.1f same %s h2d %.2f GB d2h %.2f GB" % (d["value"], d["e2e"]["value"], d["e2e"]["pageable_value"], d["e2e"]["host_equals_device_result"], d["e2e"]["h2d_bytes_per_step"]/1e9, d["e2e"]["d2h_bytes_per_step"]/1e9), d["per_rank"]["e2e_ms"], "bcast %.2f ms" % d["bvh_broadcast_ms"])
except Exception as e:
    print("$1 failed", e)
PY
}
{
run default "" "" 29531
run zerocopy1 "" "zerocopy=1" 29532
run d2h0 "" "d2h=0" 29533
run ntscatter embree-aarch64_b200/lib/variants/libembree3_nt.so "" 29534
run threads8 "" "host_threads=8" 29535
} > $OUT/${TAG}_e2e8.log 2>&1
cat $OUT/${TAG}_e2e8.log
