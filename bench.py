#!/usr/bin/env python3
"""Benchmark of the hot path: closest-hit + any-hit ray streams against a triangle BVH (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c5]

Default workload (every N): the 10 M-triangle scene of BASELINE.json configs[2]-[3] (2200x2200 displaced plane + 4
spheres); batch = 3 x (16.7 M incoherent diffuse rays, rtcIntersect1M) + 3 x (16.7 M shadow rays, rtcOccluded1M) = 100.6 M
rays, generated from a 4096x4096 primary pass with sampler seeds 0..2 / three point lights.  One step = one closest-hit
stream + one occlusion stream over the whole batch.  `value` = rays of the batch per second with the streams resident in
HBM (CUDA events on the launching stream); `e2e` = the same through the C ABI with page-locked HOST buffers (H2D, kernels,
D2H inside the timed region; a pageable-memory figure beside it).
N > 1 (one process per GPU under torchrun): rank 0 builds, the flat BVH image is broadcast over NCCL / NVLink, every rank
traces its shard (16-row bands of the frame, dealt round-robin) of the SAME batch -- strong scaling -- and the hit records are gathered on
rank 0 (timed separately: `gather`).  `--workload c2` is BASELINE configs[1] (1.0 M triangles, 33.5 M rays, L2-resident BVH),
`--workload c5` the build benchmark of configs[4].

`--impl reference` times the reference's own CPU path (oracle/_ref/libembree3_ref.so, built from the unmodified reference
sources by oracle/build_ref.py) on this box's host cores over the identical batch, driven by the pthread harness
bench/cpu_baseline.c (one thread per core, 4096-ray chunks).  That library, bench/ and oracle/ are used ONLY as baseline /
checker, never by the product.
"""
import argparse
import ctypes as C
import hashlib
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libembree3_ref.so")
CB_LIB = os.path.join(ROOT, "bench", "libcpu_baseline.so")
METRIC = "Mrays/s closest-hit & occluded (incoherent diffuse + shadow streams)"
FRAME = 4096
LIGHTS = [(5.0, 10.0, 5.0), (-5.0, 10.0, 5.0), (5.0, 10.0, -5.0), (-5.0, 10.0, -5.0), (0.0, 12.0, 0.0), (7.0, 9.0, 0.0)]
SEEDS = {"c2": 1, "c3": 3}


def load_pkg():
    pkg = importlib.import_module("embree-aarch64_b200")
    return pkg.fixtures, pkg.rtcore


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ALGORITHMIC bytes per triangle of the builder phases (default front end): what each phase must read and
# write once, not what its kernels happen to move.
BUILD_PHASE_BYTES = {"msPrims": 156, "msSort": 128, "msHierarchy": 230, "msRefit": 200, "msEmit": 160}


def build_phase_rates(build, peak_gbs):
    """Per-phase achieved GB/s (algorithmic bytes / device time) and fraction of the measured HBM peak."""
    out = {}
    n = build["numPrimsValid"]
    for k, b in BUILD_PHASE_BYTES.items():
        ms = build.get(k) or 0.0
        if ms > 0:
            gbs = n * b / (ms * 1e-3) / 1e9
            out[k] = {"ms": ms, "algorithmic_bytes_per_tri": b, "gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak_gbs}
    tot = sum(BUILD_PHASE_BYTES.values())
    out["total"] = {"ms": build["msTotal"], "algorithmic_bytes_per_tri": tot, "gb_per_s": n * tot / (build["msTotal"] * 1e-3) / 1e9,
                    "frac_of_hbm_peak": n * tot / (build["msTotal"] * 1e-3) / 1e9 / peak_gbs}
    return out


def kernel_source_hash():
    """Hash of the traversal kernel's sources: stored with an ncu capture so a stale `traffic` constant is detectable."""
    h = hashlib.sha1()
    for f in ("rq_trace.cu", "rq_math.cuh", "rq_types.h", "rq_device.h"):
        h.update(open(os.path.join(ROOT, "embree-aarch64_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.remove(self.path)
        hi = [s for s in sm if s >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def workload_meshes(fx, name):
    return fx.scene_c3(1.0) if name == "c3" else fx.scene_c2(1.0)


def workload_name(w, tris, rays):
    if w == "c3":
        return (f"configs[2]-[3]: {tris}-triangle displaced plane + 4 spheres, 4096x4096 primary pass -> {rays} incoherent rays per step "
                f"({SEEDS[w]} diffuse closest-hit streams (rtcIntersect1M) + {SEEDS[w]} shadow streams (rtcOccluded1M), sampler seeds 0..{SEEDS[w] - 1})")
    return (f"configs[1]: {tris}-triangle displaced plane + sphere, 4096x4096 primary pass -> {rays} incoherent rays per step "
            f"(diffuse closest-hit stream (rtcIntersect1M) + shadow stream (rtcOccluded1M))")


SHARD_ROWS = 16                                                            # N > 1: frame rows are dealt to the ranks in bands of this many rows


def shard_bands(rank, world):
    """Row bands [(r0, r1), ...] of the frame that rank `rank` of `world` traces.  One GPU: the whole frame (in 8 pieces, to bound the
    host memory of the generator).  N GPUs: 16-row bands dealt round-robin -- contiguous N-ths of this frame differ by 10 % in
    cost per ray (max / mean 1.106 at N = 8, 1.026 with 64-row bands; profiles/r02l_shard_balance.jsonl), and strong scaling is
    timed as the max over ranks."""
    if world == 1:
        return [(FRAME * b // 8, FRAME * (b + 1) // 8) for b in range(8)]
    return [(r0, min(r0 + SHARD_ROWS, FRAME)) for k, r0 in enumerate(range(0, FRAME, SHARD_ROWS)) if k % world == rank]


def make_streams(fx, trace_primary, bands, seeds):
    """The batch of this process: for every sampler seed one diffuse and one shadow stream over the frame rows of `bands`
    (list of (r0, r1)).  trace_primary(rays) traces a coherent primary stream in place.  Returns (diffuse RAYHIT array, shadow RAY array)."""
    d_parts = [[] for _ in range(seeds)]
    s_parts = [[] for _ in range(seeds)]
    for r0, r1 in bands:
        if r1 <= r0:
            continue
        prim = fx.primary_rays(FRAME, FRAME, rows=(r0, r1), **fx.C2_CAMERA)
        trace_primary(prim)
        for k in range(seeds):
            d_parts[k].append(fx.diffuse_rays(prim, sample_id=k))
            s_parts[k].append(fx.shadow_rays(prim, light=LIGHTS[k]))
    diffuse = np.concatenate([p for k in range(seeds) for p in d_parts[k]])
    shadow = np.concatenate([p for k in range(seeds) for p in s_parts[k]])
    return diffuse, shadow


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU path driven by bench/cpu_baseline.c
# ------------------------------------------------------------------------------------------------
class CpuDriver:
    def __init__(self, ref):
        if not os.path.exists(CB_LIB):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "bench")])
        self.cb = C.CDLL(CB_LIB)
        self.cb.cb_trace_stream.restype = C.c_double
        self.cb.cb_trace_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_int]
        self.f_int = C.cast(ref.lib.rtcIntersect1M, C.c_void_p).value
        self.f_occ = C.cast(ref.lib.rtcOccluded1M, C.c_void_p).value
        self.cores = os.cpu_count() or 1

    def trace(self, scene, rays, occluded=False, coherent=False, chunk=4096):
        t = self.cb.cb_trace_stream(self.f_occ if occluded else self.f_int, scene, 1 if coherent else 0, rays.ctypes.data, len(rays),
                                    rays.strides[0], chunk, self.cores)
        if t < 0:
            raise RuntimeError("cpu_baseline driver failed")
        return t


def reference_setup(fx, rt, workload, seeds, streams=None):
    """streams: (diffuse, shadow) to reuse instead of generating them with the reference library (the cpu_baseline leg of our
    arm hands over the seed-0 streams it already has; the reference arm proper generates its own)."""
    if not os.path.exists(REF_LIB):
        return None
    ref = rt.RTCore(REF_LIB)
    drv = CpuDriver(ref)
    dev = ref.new_device(f"threads={drv.cores}")
    meshes = workload_meshes(fx, workload)
    t0 = time.perf_counter()
    sc, keep = ref.build_scene(dev, meshes)
    build_s = time.perf_counter() - t0
    diffuse, shadow = streams if streams is not None else make_streams(fx, lambda r: drv.trace(sc, r, coherent=True), shard_bands(0, 1), seeds)
    return dict(ref=ref, drv=drv, dev=dev, sc=sc, keep=keep, cores=drv.cores, diffuse=diffuse, shadow=shadow, build_s=build_s, tris=fx.num_tris(meshes))


def reference_step(S):
    d, s = S["diffuse"].copy(), S["shadow"].copy()                          # fresh inputs every step, not timed
    t = S["drv"].trace(S["sc"], d)
    t += S["drv"].trace(S["sc"], s, occluded=True)
    return t, len(d) + len(s)


def run_reference(args):
    fx, rt = load_pkg()
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    S = reference_setup(fx, rt, args.workload, SEEDS[args.workload])
    if S is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libembree3_ref.so not built (run oracle/build_ref.py where /root/reference exists)"}))
        return 0
    for _ in range(args.warmup):
        reference_step(S)
    tot_t, tot_n = 0.0, 0
    for _ in range(args.steps):
        t, n = reference_step(S)
        tot_t += t; tot_n += n
    v = tot_n / tot_t / 1e6
    nrays = len(S["diffuse"]) + len(S["shadow"])
    sample = (f"the full batch every step: {len(S['diffuse'])} diffuse + {len(S['shadow'])} shadow rays, one pthread per host core, "
              f"4096-ray rtcIntersect1M / rtcOccluded1M calls (bench/cpu_baseline.c)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, S["tris"], nrays), "rays_per_step": nrays, "same_config": True},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": S["cores"], "kind": "reference", "sample": sample,
                             "build_mtris_per_s": S["tris"] / S["build_s"] / 1e6},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def l2_copy_bandwidth(torch):
    """Measured L2 peak for L2-resident workloads (configs[0]-[1]): the better of (a) a device copy between two 24 MiB buffers and
    (b) a read-only reduction over 64 MiB, both resident in the 126 MB L2 after the first pass; timed in batches so that launch
    latency does not dominate the ~10 us kernels."""
    def timed(fn, nbytes, reps=32):
        for _ in range(4):
            fn()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        return nbytes / (best * 1e-3) / 1e9
    a = torch.empty(24 << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    copy = timed(lambda: b.copy_(a), 2 * a.numel())
    c = torch.zeros(16 << 20, dtype=torch.float32, device="cuda")
    read = timed(lambda: c.sum(), 4 * c.numel())
    return max(copy, read)


def run_ours(args):
    import torch
    import torch.distributed as dist
    fx, rt = load_pkg()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = rt.RTCore(os.environ["RQ_B200_LIB"]) if os.environ.get("RQ_B200_LIB") else rt.RTCore()   # fails loudly when the CUDA library is missing (RQ_B200_LIB: experiment build)
    threads = max(2, min(8, (os.cpu_count() or 8) // max(world, 1)))       # host staging threads per process: N ranks share the box's cores
    dev = lib.new_device(f"gpu={local},async=1,host_threads={threads}" + ("," + os.environ["RQ_B200_CFG"] if os.environ.get("RQ_B200_CFG") else ""))   # RQ_B200_CFG: extra device options for experiments
    # a real (non-default) stream shared by torch and the library: the CUDA events below are recorded on
    # the stream the kernels are launched on (handle 0 would mean "the library's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    lib.lib.rtcxSetDeviceStream(dev, stream.cuda_stream)
    meshes = workload_meshes(fx, args.workload)
    ntris = fx.num_tris(meshes)
    seeds = SEEDS[args.workload]

    # ---- build on rank 0, replicate the flat image over NCCL (collective 1 of 2) ----
    bcast_ms, build = 0.0, None
    if rank == 0:
        sc, keep = lib.build_scene(dev, meshes)
        build_times = []
        for _ in range(3):                                                 # rtcCommitScene wall time, re-committed (buildbench style)
            for g in range(len(meshes)):
                lib.lib.rtcCommitGeometry(lib.lib.rtcGetGeometry(sc, g))
            t0 = time.perf_counter(); lib.lib.rtcCommitScene(sc); build_times.append(time.perf_counter() - t0)
        build = lib.build_stats(sc)
        build["commit_wall_ms"] = float(np.median(build_times) * 1e3)
    if world > 1:
        mg = importlib.import_module("embree-aarch64_b200.multigpu")
        sc, bcast_ms = mg.replicate_scene(lib, dev, sc if rank == 0 else None, 0)
    assert lib.lib.rtcGetDeviceError(dev) == 0

    # ---- this rank's shard of the batch: its row bands of every stream (the whole frame on one GPU) ----
    diffuse, shadow = make_streams(fx, lambda r: lib.intersect(sc, r, coherent=True), shard_bands(rank, world), seeds)
    nd, ns = len(diffuse), len(shadow)
    h_d = torch.from_numpy(diffuse.view(np.uint8).reshape(nd, 80)).pin_memory()
    h_s = torch.from_numpy(shadow.view(np.uint8).reshape(ns, 48)).pin_memory()
    diffuse = h_d.numpy().reshape(-1).view(rt.RAYHIT_DTYPE)                 # keep one host copy of the pristine batch (the page-locked one)
    shadow = h_s.numpy().reshape(-1).view(rt.RAY_DTYPE)
    p_d, p_s = h_d.cuda(), h_s.cuda()                                       # pristine device copies
    w_d, w_s = torch.empty_like(p_d), torch.empty_like(p_s)                 # working copies (traced in place)

    # ---- counters (instrumented kernels, outside the timed region): the roofline numerator ----
    w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
    c_close = lib.intersect_counted(sc, w_d.data_ptr(), nd, 80)
    c_occ = lib.intersect_counted(sc, w_s.data_ptr(), ns, 48, occluded=True)
    hits = int(c_close["hits"])

    # ---- timed region: device-resident streams, CUDA events on the launching stream ----
    def step_device():
        lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80)
        lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)
    for _ in range(args.warmup):
        w_d.copy_(p_d); w_s.copy_(p_s); step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = lib.lib.rtcxGetLaunchCount()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(args.steps)]
    for k in range(args.steps):
        w_d.copy_(p_d); w_s.copy_(p_s)                                      # fresh inputs (also flushes L2: GBs of rays >> 126 MB); not timed
        ev[k][0].record(); lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80)
        ev[k][1].record(); lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)
        ev[k][2].record()
    torch.cuda.synchronize()
    launches = lib.lib.rtcxGetLaunchCount() - launches0
    if world > 1:
        dist.barrier()
    t_close = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    t_occ = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    ms_step = t_close + t_occ

    # ---- collective 2 of 2: the hit records (tfar + hit, bytes 32..79 of every closest-hit record) gathered on rank 0 ----
    gather_ms, gather_bytes = 0.0, 0
    if world > 1:
        hitpart = torch.empty((nd, 48), dtype=torch.uint8, device="cuda")
        out = [torch.empty_like(hitpart) for _ in range(world)] if rank == 0 else None
        for k in range(3):
            torch.cuda.synchronize(); dist.barrier()
            g0, g1 = torch.cuda.Event(True), torch.cuda.Event(True)
            g0.record()
            hitpart.copy_(w_d[:, 32:])
            dist.gather(hitpart, out, 0)
            g1.record(); torch.cuda.synchronize()
            gather_ms = g0.elapsed_time(g1)
        gather_bytes = nd * 48 * (world - 1)
        del hitpart, out

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H timed) ----
    e2e = {}
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))
        hw_d, hw_s = torch.empty_like(h_d).pin_memory(), torch.empty_like(h_s).pin_memory()
        t_e2e = 0.0
        xfer0 = (0, 0)
        for k in range(1 + e2e_steps):
            if k == 1:
                xfer0 = lib.transfer_bytes(dev)                            # bytes the library itself copies over PCIe (counted at its cudaMemcpy calls)
            hw_d.copy_(h_d); hw_s.copy_(h_s)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            lib.intersect_ptr(sc, hw_d.data_ptr(), nd, 80)                  # host pointer: staged by the library
            lib.occluded_ptr(sc, hw_s.data_ptr(), ns, 48)
            dt = time.perf_counter() - t0
            if k > 0:
                t_e2e += dt
        t_e2e /= e2e_steps
        xfer1 = lib.transfer_bytes(dev)
        h2d_step, d2h_step = (xfer1[0] - xfer0[0]) // e2e_steps, (xfer1[1] - xfer0[1]) // e2e_steps
        same = bool(np.array_equal(hw_d.numpy(), w_d.cpu().numpy()))         # host path == device path, bit for bit
        # the route a drop-in application takes: malloc'ed (pageable) ray buffers handed to rtcIntersect1M / rtcOccluded1M
        pg_d, pg_s = diffuse.copy(), shadow.copy()
        t_pg = 0.0
        for k in range(2):
            pg_d[:] = diffuse; pg_s[:] = shadow
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            lib.intersect_ptr(sc, pg_d.ctypes.data, nd, 80)
            lib.occluded_ptr(sc, pg_s.ctypes.data, ns, 48)
            t_pg = time.perf_counter() - t0                                 # second pass: staging buffers exist
        same = same and bool(np.array_equal(pg_d.view(np.uint8).reshape(nd, 80), hw_d.numpy()))
        e2e = dict(t=t_e2e, t_pg=t_pg, h2d=h2d_step, d2h=d2h_step, same=same)
    clocks = sampler.stop()

    # ---- max over ranks ----
    per_rank = None
    if world > 1:
        t = torch.tensor([ms_step, e2e.get("t", 0.0) * 1e3, t_close, t_occ, e2e.get("t_pg", 0.0) * 1e3, gather_ms], dtype=torch.float64, device="cuda")
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)                                            # every rank's own times: how even the shards are
        per_rank = {"device_ms": [round(float(x[0]), 3) for x in allt], "e2e_ms": [round(float(x[1]), 3) for x in allt]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, t_close, t_occ, pg_ms, gather_ms = [float(x) for x in t.tolist()]
        cnt = torch.tensor([nd + ns, launches, e2e.get("h2d", 0), e2e.get("d2h", 0), nd, ns, c_close["nodes"], c_close["tris"], hits,
                            c_occ["nodes"], c_occ["tris"], c_occ["rays"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        total_rays, launches, h2d_step, d2h_step, tnd, tns, cn, ct, hits, on, ot, orays = [int(x) for x in cnt.tolist()]
    else:
        e2e_ms, pg_ms, total_rays = e2e.get("t", 0.0) * 1e3, e2e.get("t_pg", 0.0) * 1e3, nd + ns
        h2d_step, d2h_step, tnd, tns = e2e.get("h2d", 0), e2e.get("d2h", 0), nd, ns
        cn, ct, on, ot, orays = c_close["nodes"], c_close["tris"], c_occ["nodes"], c_occ["tris"], c_occ["rays"]

    if rank == 0:
        peaks = measured_peaks()
        node_b, tri_b = 80, 48                                             # bytes of a node / triangle record traversal reads
        # ALGORITHMIC bytes of the dominant kernel (closest hit) over all ranks / its launch time (max over ranks)
        alg_close = cn * node_b + ct * tri_b + tnd * 48 + hits * 36
        ach = alg_close / (t_close * 1e-3) / 1e9 / world                   # per GPU
        image_mb = (build["bytes"] / 1e6) if build else None
        l2_mb = torch.cuda.get_device_properties(local).L2_cache_size / 1e6
        l2_peak = None                                                      # measured only for L2-resident workloads (two torch kernels: kept out of the default run)
        if image_mb is not None and image_mb < 0.75 * l2_mb:
            l2_peak = l2_copy_bandwidth(torch)
            bound, peak = "l2", l2_peak
            peak_src = f"measured in this run: best of an L2-resident 24 MiB device copy (read+write) and a 64 MiB read-only reduction; the {image_mb:.0f} MB BVH image fits the {l2_mb:.0f} MB L2"
        else:
            bound = "hbm"
            peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
            peak_src = ("MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)") + \
                       f"; the {image_mb:.0f} MB BVH image does not fit the {l2_mb:.0f} MB L2"
        traffic, traffic_src = None, None
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
            if ncu and world == 1:
                stale = ncu.get("kernel_source_hash") != kernel_source_hash()
                scale = nd / ncu["rays_per_launch"]                        # same stream, possibly captured on a shorter launch
                traffic = None if stale else ncu["dram_bytes_per_launch_closest"] * scale
                traffic_src = {"file": "profiles/ncu_traffic.json", "capture": ncu.get("capture"), "kernel_source_hash": ncu.get("kernel_source_hash"),
                               "stale": stale, "lts_hit_rate_pct": ncu.get("lts_hit_rate_pct"), "lanes_per_inst": ncu.get("lanes_per_inst")}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, ntris, total_rays), "rays_per_step": total_rays, "rays_per_gpu_per_step": nd + ns,
                       "l2_policy": "inputs (GBs of ray records per step) exceed the 126 MB L2; streams re-copied from pristine buffers between timed steps",
                       "parallelism": (f"frame rows dealt to the {world} ranks in {SHARD_ROWS}-row bands (strong scaling: the batch is fixed), BVH replica per GPU"
                                       if world > 1 else "one GPU traces the whole batch")},
            "closest_mrays_per_s": tnd / (t_close * 1e-3) / 1e6, "occluded_mrays_per_s": tns / (t_occ * 1e-3) / 1e6,
            "build": build, "build_mtris_per_s": (ntris / (build["msTotal"] * 1e-3) / 1e6) if build else None,
            "build_phases": build_phase_rates(build, float(peaks["hbm_gbs"]) if peaks else 6650.0) if build else None,
            "bvh_broadcast_ms": bcast_ms,
            "gather": {"ms": gather_ms, "bytes": gather_bytes, "what": "48 B (tfar + hit) of every closest-hit record, NCCL gather to rank 0",
                       "value_with_gather": total_rays / ((ms_step + gather_ms) * 1e-3) / 1e6} if world > 1 else None,
            "traversal_per_ray": {"closest_nodes": cn / tnd, "closest_tris": ct / tnd,
                                  "occluded_nodes": on / max(orays, 1), "occluded_tris": ot / max(orays, 1), "hit_fraction": hits / tnd},
            "roofline": {"bound": bound, "kernel": "k_trace<closest>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_source": peak_src, "bytes_per_ray": alg_close / tnd,
                         "achieved_if_nodes_count_128B": (alg_close + cn * 48) / (t_close * 1e-3) / 1e9 / world,
                         "launch_ms": t_close, "rays_per_launch": nd, "traffic": traffic, "traffic_source": traffic_src,
                         "frac_of_hbm_peak": ach / (float(peaks["hbm_gbs"]) if peaks else 6650.0),
                         "frac_of_measured_l2_peak": (ach / l2_peak) if l2_peak else None, "measured_l2_peak_gbs": l2_peak,
                         "note": "achieved = ALGORITHMIC bytes (80 B per node record + 48 B per triangle record fetched, counted by the instrumented kernel, "
                                 "+ 48 B per ray in + 36 B per hit out) / launch time. `traffic` = DRAM bytes of the same launch from the committed ncu capture: "
                                 "well below the algorithmic bytes because L1 / L2 serve the upper levels of the tree (see traffic_source: L2 hit rate, "
                                 "active lanes per instruction) -- the kernel is issue / divergence bound, not DRAM bound, on both scenes"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if per_rank:
            line["per_rank"] = per_rank
        if e2e:
            line["e2e"] = {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d_step),
                           "d2h_bytes_per_step": int(d2h_step), "host_record_bytes_per_step": tnd * 80 + tns * 48,
                           "pageable_value": total_rays / (pg_ms * 1e-3) / 1e6,
                           "path": "rtcIntersect1M + rtcOccluded1M on page-locked host streams: H2D of the ray records, kernels, compact "
                                   "hit-list D2H, scatter into the caller's records by the library's host threads (all inside the timed region); "
                                   "pageable_value = the same calls on malloc'ed buffers",
                           "host_equals_device_result": e2e["same"]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                S = reference_setup(fx, rt, args.workload, 1, streams=(diffuse[:nd // seeds].copy(), shadow[:ns // seeds].copy()))   # bounded sample: the seed-0 streams of the batch
                if S is not None:
                    reference_step(S)
                    tt, nn = 0.0, 0
                    for _ in range(3):
                        t, n = reference_step(S)
                        tt += t; nn += n
                    line["cpu_baseline"] = {"value": nn / tt / 1e6, "unit": "Mrays/s", "cores": S["cores"], "kind": "reference",
                                            "sample": f"3 x ({len(S['diffuse'])} diffuse + {len(S['shadow'])} shadow rays: the seed-0 streams of the batch), "
                                                      "one pthread per host core, 4096-ray calls (bench/cpu_baseline.c)",
                                            "build_mtris_per_s": S["tris"] / S["build_s"] / 1e6}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
            except Exception as e:                                          # the baseline must never take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_build_bench(args):
    """BASELINE.json configs[4]: builds of 10 / 20 / 50 M triangles (+ the unstructured soup) with the reference's builder beside them."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    bb = importlib.import_module("bench_build")
    rows = bb.run(sizes=[float(x) for x in args.sizes.split(",")], kinds=args.kinds.split(","), reference=not args.no_cpu_baseline)
    peaks = measured_peaks()
    main_row = next((r for r in rows if r["kind"] == "scene"), rows[0])
    line = {"metric": "BVH build Mtris/s (device time, rtcCommitScene)", "value": main_row["ours"]["mtris_per_s_device"], "unit": "Mtris/s",
            "n_gpus": 1, "steps": 3, "warmup": 1, "ms_per_step": main_row["ours"]["device_ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: GPU BVH build of 10-50 M triangles vs the reference's binned-SAH BVH8 builder, SAH + probe-stream Mrays/s on each tree"},
            "hbm_peak_gbs": float(peaks["hbm_gbs"]) if peaks else None, "builds": rows, "gpu_launches": int(bb.LAUNCHES[0])}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sizes", default="10,20,50")
    ap.add_argument("--kinds", default="scene,soup")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c5":
        return run_build_bench(args)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
