"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo group on 127.0.0.1.
The GPU variant (NCCL broadcast of a real BVH image, N=2) is test_gpu_multi below, marked gpu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

mg = cases.importlib.import_module("embree-aarch64_b200.multigpu")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. image broadcast: only rank 0 knows the payload
        payload = torch.arange(100003, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.empty(0, dtype=torch.uint8)
        got = mg.broadcast_bytes(payload, 0)
        ok = got.numel() == 100003 and int(got.to(torch.int64).sum()) == int(torch.arange(100003).to(torch.uint8).to(torch.int64).sum())
        # 2. ray sharding + gather of hit slices: every ray exactly once, order preserved
        M = 1000 * world + 7
        b, e = mg.shard_range(M, rank, world)
        local = (torch.arange(b, e, dtype=torch.int64) % 251).to(torch.uint8)
        counts = [mg.shard_range(M, r, world)[1] - mg.shard_range(M, r, world)[0] for r in range(world)]
        parts = mg.gather_slices(local, counts, 0)
        if rank == 0:
            whole = torch.cat(parts)
            ok = ok and whole.numel() == M and bool((whole == (torch.arange(M) % 251).to(torch.uint8)).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_everything():
    for M in (0, 1, 7, 100, 16777211):
        for w in (1, 2, 3, 8):
            r = [mg.shard_range(M, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_gloo_world2_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert sorted(res) == [(0, True), (1, True)]
