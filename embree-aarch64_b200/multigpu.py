"""Multi-GPU plumbing: one process per GPU (torch.distributed), ray streams sharded, BVH replicated.

The path shards naturally -- rays are independent and the committed BVH is read-only --
so there are exactly two collectives and neither is on the traversal data path:
  1. broadcast of the flat BVH image from the building rank (NCCL over NVLink / NVSwitch), after
     which every rank adopts its byte copy with rtcxSetSceneImage;
  2. (optional) gather of per-rank hit slices when the caller wants the whole stream on one rank.
torch is used for device memory and the process group only.
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_range(num_rays, rank, world):
    """Contiguous range [begin, end) of rank `rank`: keeps whatever coherence the stream has."""
    base, rem = divmod(int(num_rays), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_bytes(payload, src=0, device="cpu"):
    """Broadcast a uint8 tensor whose size only `src` knows.  Returns the tensor on every rank."""
    rank = dist.get_rank()
    n = torch.tensor([payload.numel() if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    buf = payload if rank == src else torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf


def replicate_scene(lib, device, scene, src=0):
    """Rank `src` holds a committed `scene`; every other rank passes scene=None and receives a
    replica.  Returns (scene, broadcast_ms).  Image bytes travel GPU-to-GPU (NCCL)."""
    rank = dist.get_rank()
    if rank == src:
        nbytes = C.c_size_t(0)
        lib.lib.rtcxGetSceneImage(scene, C.byref(nbytes))
        img = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
        lib.lib.rtcxCopySceneImage(scene, img.data_ptr(), nbytes.value)
    else:
        img = torch.empty(0, dtype=torch.uint8, device="cuda")
    # NCCL sets its broadcast channels (and, per message size class, its protocol buffers) up lazily: a one-word and a 32 MB
    # broadcast first, so that the time reported is the transfer
    for n in (1, 32 << 20):
        warm = torch.zeros(n, dtype=torch.uint8, device="cuda")
        dist.broadcast(warm, src)
    del warm
    # size first, receive buffers allocated before the clock starts (a fresh 0.66 GB cudaMalloc in eight processes at once took
    # tens of ms and was being reported as transfer time: 46 ms at N = 8 against 2.2 ms at N = 4, profiles/r02t_bench_8gpu.json)
    n = torch.tensor([img.numel() if rank == src else 0], dtype=torch.int64, device="cuda")
    dist.broadcast(n, src)
    if rank != src:
        img = torch.empty(int(n.item()), dtype=torch.uint8, device="cuda")
        img.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.broadcast(img, src)
    e1.record()
    torch.cuda.synchronize()
    if rank != src:
        scene = lib.lib.rtcNewScene(device)
        lib.lib.rtcxSetSceneImage(scene, img.data_ptr(), img.numel())
        if lib.lib.rtcGetDeviceError(device) != 0:
            raise RuntimeError("replica import failed")
    return scene, e0.elapsed_time(e1)


def gather_slices(local, counts, dst=0):
    """Gather variable-length uint8 slices (hit records) on rank `dst`; returns the list there, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    m = max(counts)
    pad = torch.zeros(m, dtype=torch.uint8, device=local.device)
    pad[:local.numel()] = local
    out = [torch.empty(m, dtype=torch.uint8, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst)
    return [o[:c] for o, c in zip(out, counts)] if rank == dst else None
