#!/bin/bash
# 2-GPU box: gpus=2 through the C ABI (tests + end-to-end), bench.py --gpus 2 under torchrun; plus the single-GPU staged-upload test and commit wall times
OUT=gpurun_out; TAG=${1:-r02r}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1; nproc >> $OUT/${TAG}_topo.txt
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_parity.py -m gpu -q -k "gpus2 or multi or staged or full_size" > $OUT/${TAG}_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_2gpu.log; tail -3 $OUT/${TAG}_pytest_2gpu.log
python tools/profile_build.py --workload c3 > $OUT/${TAG}_build_c3_staged.jsonl 2>&1
python tools/profile_build.py --workload c3 --cfg stage_geometry=0 > $OUT/${TAG}_build_c3_plain.jsonl 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
tail -c 1500 $OUT/${TAG}_bench_2gpu.json
timeout 900 python tools/bench_cabi_gpus.py > $OUT/${TAG}_cabi_gpus.jsonl 2> $OUT/${TAG}_cabi_gpus.err; cat $OUT/${TAG}_cabi_gpus.jsonl
