#!/bin/bash
# 8-GPU box: the strong-scaling bench at N = 8 (full line incl. e2e) and N = 4 (device-timed only), launched like the driver does
OUT=gpurun_out; TAG=${1:-r02t}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo8.txt 2>&1; nproc >> $OUT/${TAG}_topo8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/${TAG}_bench_8gpu.json 2> $OUT/${TAG}_bench_8gpu.err
tail -c 1200 $OUT/${TAG}_bench_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e > $OUT/${TAG}_bench_4gpu.json 2> $OUT/${TAG}_bench_4gpu.err
tail -c 600 $OUT/${TAG}_bench_4gpu.json
