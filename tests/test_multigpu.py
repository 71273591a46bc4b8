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


# ------------------------------------------------------------------------------------------------------------------
# gpus=N behind the C ABI (one process, N GPUs): needs >= 2 devices -- run with `gpurun --gpus 2`; skipped on a 1-GPU box
# ------------------------------------------------------------------------------------------------------------------
def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")


@pytest.mark.gpu
def test_gpus2_sharded_host_stream_is_bit_identical(product):
    """rtcNewDevice("gpus=2"): the commit replicates the image to GPU 1 (cudaMemcpyPeer), one rtcIntersect1M / rtcOccluded1M
    call on a host stream is sharded over both GPUs, and the caller's buffer ends up exactly as a single GPU leaves it."""
    _need_two_gpus()
    fx, rt = cases.fx, cases.rt
    meshes = fx.scene_c2(0.3)
    one = product.new_device("gpu=0")
    two = product.new_device("gpu=0,gpus=2,shard_min_rays=4096")
    assert product.lib.rtcxGetDeviceGpuCount(one) == 1 and product.lib.rtcxGetDeviceGpuCount(two) == 2
    sc1, k1 = product.build_scene(one, meshes)
    sc2, k2 = product.build_scene(two, meshes)
    st = product.build_stats(sc2)
    assert st["msBroadcast"] > 0.0
    prim = fx.primary_rays(512, 512, **fx.C2_CAMERA)
    product.intersect(sc1, prim, coherent=True)
    d = fx.diffuse_rays(prim)
    a, b = d.copy(), d.copy()
    x0 = product.transfer_bytes(two)
    product.intersect(sc1, a)
    product.intersect(sc2, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert product.transfer_bytes(two)[0] - x0[0] >= len(d) * 32          # both shards went over PCIe through the library
    s = fx.shadow_rays(prim)
    a, b = s.copy(), s.copy()
    product.occluded(sc1, a)
    product.occluded(sc2, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    # a stream resident on GPU 1 is traced there, against that GPU's replica
    want = d.copy()
    product.intersect(sc1, want)
    t1 = torch.from_numpy(d.view(np.uint8).reshape(len(d), 80).copy()).to("cuda:1")
    product.intersect_ptr(sc2, t1.data_ptr(), len(d))
    torch.cuda.synchronize(1)
    assert np.array_equal(t1.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE), want)
    # re-commit after a vertex update: the replica follows
    k2[0][:meshes[0][0].size] += np.float32(0.01)
    g0 = product.lib.rtcGetGeometry(sc2, 0)
    product.lib.rtcUpdateGeometryBuffer(g0, rt.RTC_BUFFER_TYPE_VERTEX, 0); product.lib.rtcCommitGeometry(g0); product.lib.rtcCommitScene(sc2)
    k1[0][:meshes[0][0].size] += np.float32(0.01)
    g0 = product.lib.rtcGetGeometry(sc1, 0)
    product.lib.rtcUpdateGeometryBuffer(g0, rt.RTC_BUFFER_TYPE_VERTEX, 0); product.lib.rtcCommitGeometry(g0); product.lib.rtcCommitScene(sc1)
    a, b = d.copy(), d.copy()
    product.intersect(sc1, a)
    product.intersect(sc2, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert product.lib.rtcGetDeviceError(two) == 0 and product.lib.rtcGetDeviceError(one) == 0
    for sc in (sc1, sc2):
        product.lib.rtcReleaseScene(sc)
    product.lib.rtcReleaseDevice(one); product.lib.rtcReleaseDevice(two)


@pytest.mark.gpu
def test_nccl_replica_on_second_gpu_answers_identically(product):
    """The torchrun path's replica (rtcxCopySceneImage -> bytes -> rtcxSetSceneImage on another device, validated on adoption)
    answers bit-identically on device 1."""
    _need_two_gpus()
    import ctypes as C
    fx, rt = cases.fx, cases.rt
    meshes = fx.scene_c2(0.3)
    d0 = product.new_device("gpu=0")
    d1 = product.new_device("gpu=1")
    sc0, keep = product.build_scene(d0, meshes)
    n = C.c_size_t(0)
    product.lib.rtcxGetSceneImage(sc0, C.byref(n))
    img = torch.empty(n.value, dtype=torch.uint8, device="cuda:0")
    product.lib.rtcxCopySceneImage(sc0, img.data_ptr(), n.value)
    img1 = img.to("cuda:1")
    sc1 = product.lib.rtcNewScene(d1)
    product.lib.rtcxSetSceneImage(sc1, img1.data_ptr(), n.value)
    assert product.lib.rtcGetDeviceError(d1) == 0
    r = fx.incoherent_rays(200000, org=(0.0, 3.0, 0.0), seed=3)
    a, b = r.copy(), r.copy()
    product.intersect(sc0, a)
    product.intersect(sc1, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and (a["geomID"] != 0xFFFFFFFF).any()
    product.lib.rtcReleaseScene(sc0); product.lib.rtcReleaseScene(sc1)
    product.lib.rtcReleaseDevice(d0); product.lib.rtcReleaseDevice(d1)


def test_bench_shard_bands_cover_the_frame_once():
    """bench.py deals the frame rows to the ranks in 16-row bands: every row exactly once, equal shares, one GPU = the whole frame."""
    import bench
    for w in (1, 2, 3, 4, 8):
        bands = [bench.shard_bands(r, w) for r in range(w)]
        rows = sorted(x for b in bands for (a, c) in b for x in range(a, c))
        assert rows == list(range(bench.FRAME))
        share = [sum(c - a for a, c in b) for b in bands]
        assert max(share) - min(share) <= bench.SHARD_ROWS
