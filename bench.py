#!/usr/bin/env python3
"""Benchmark of the hot path: closest-hit + any-hit ray streams against a triangle BVH (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Workload at N=1 (BASELINE.json configs[1]): ~1.0 M-triangle displaced plane + sphere, 4096x4096 primary
pass -> 16.7 M incoherent diffuse rays (rtcIntersect1M) + 16.7 M shadow rays (rtcOccluded1M).
One step = one closest-hit stream + one occlusion stream over the whole batch; `value` is rays of
both streams per second with the streams resident in HBM; `e2e` is the same through the C ABI with
pinned HOST buffers (H2D + kernels + D2H inside the timed region).  N>1: one process per GPU
(torchrun), rank 0 builds, the flat BVH image is broadcast over NCCL/NVLink, every rank traces its
own shard of rays (weak scaling: per-GPU batch fixed), no collective on the data path.

`--impl reference` times the reference's own CPU path (oracle/_ref/libembree3_ref.so, built from the
unmodified reference sources by oracle/build_ref.py) on this box's host cores on a bounded sample of
the same workload.  That library and oracle/ are used ONLY as baseline/checker, never by the product.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libembree3_ref.so")
METRIC = "Mrays/s closest-hit & occluded (incoherent diffuse + shadow streams)"


def load_pkg():
    pkg = importlib.import_module("embree-aarch64_b200")
    return pkg.fixtures, pkg.rtcore


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


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


def camera(fx):
    return fx.C2_CAMERA


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU path, all host threads, bounded sample
# ------------------------------------------------------------------------------------------------
def threaded_stream(lib_call, rays, nthreads, chunk=4096):
    """One std::thread-like worker per hardware thread, each looping over 4096-ray chunks.
    ctypes releases the GIL inside the library call, so the workers run in parallel."""
    n = len(rays)
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                s = nxt[0]
                nxt[0] += chunk * 16
            if s >= n:
                return
            for c in range(s, min(n, s + chunk * 16), chunk):
                lib_call(rays[c:c + chunk])
    th = [threading.Thread(target=work) for _ in range(nthreads)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    return time.perf_counter() - t0


def reference_setup(fx, rt, workload, rows):
    if not os.path.exists(REF_LIB):
        return None
    ref = rt.RTCore(REF_LIB)
    cores = os.cpu_count() or 1
    dev = ref.new_device(f"threads={cores}")
    meshes = workload_meshes(fx, workload)
    t0 = time.perf_counter()
    sc, keep = ref.build_scene(dev, meshes)
    build_s = time.perf_counter() - t0
    prim = fx.primary_rays(4096, 4096, rows=rows, **camera(fx))
    threaded_stream(lambda r: ref.intersect(sc, r, coherent=True), prim, cores)
    diffuse = fx.diffuse_rays(prim)
    probe = diffuse.copy()
    threaded_stream(lambda r: ref.intersect(sc, r), probe, cores)
    shadow = fx.shadow_rays(prim)
    return dict(ref=ref, dev=dev, sc=sc, keep=keep, cores=cores, diffuse=diffuse, shadow=shadow, build_s=build_s, tris=fx.num_tris(meshes))


def reference_step(S):
    d, s = S["diffuse"].copy(), S["shadow"].copy()
    t = threaded_stream(lambda r: S["ref"].intersect(S["sc"], r), d, S["cores"])
    t += threaded_stream(lambda r: S["ref"].occluded(S["sc"], r), s, S["cores"])
    return t, len(d) + len(s)


def run_reference(args):
    fx, rt = load_pkg()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    S = reference_setup(fx, rt, args.workload, rows=(1792, 2304))          # 512 rows of the 4096x4096 frame = 2.1 M rays per stream
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
    sample = f"{len(S['diffuse'])} diffuse + {len(S['shadow'])} shadow rays per step (rows 1792-2303 of the 4096x4096 frame), 4096-ray chunks per thread"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, S["tris"]), "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": S["cores"], "kind": "reference", "sample": sample,
                             "build_mtris_per_s": S["tris"] / S["build_s"] / 1e6},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_name(w, tris):
    if w == "c3":
        return f"configs[2]: {tris}-triangle displaced plane + 4 spheres, 4096x4096 primary -> incoherent diffuse + shadow streams"
    return f"configs[1]: {tris}-triangle displaced plane + sphere, 4096x4096 primary -> 16.7M incoherent diffuse rays (rtcIntersect1M) + shadow rays (rtcOccluded1M)"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
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
    dev = lib.new_device(f"gpu={local},async=1" + ("," + os.environ["RQ_B200_CFG"] if os.environ.get("RQ_B200_CFG") else ""))   # RQ_B200_CFG: extra device options for experiments
    # a real (non-default) stream shared by torch and the library: the CUDA events below are recorded on
    # the stream the kernels are launched on (handle 0 would mean "the library's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    lib.lib.rtcxSetDeviceStream(dev, stream.cuda_stream)
    meshes = workload_meshes(fx, args.workload)
    ntris = fx.num_tris(meshes)

    # ---- build on rank 0, replicate the flat image over NCCL (the only collective besides the timing all-reduce) ----
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

    # ---- this rank's shard: full 4096x4096 frame, sampler seed = rank (weak scaling) ----
    bands = 8
    d_parts, s_parts = [], []
    for b in range(bands):
        prim = fx.primary_rays(4096, 4096, rows=(b * 4096 // bands, (b + 1) * 4096 // bands), **camera(fx))
        lib.intersect(sc, prim, coherent=True)
        d_parts.append(fx.diffuse_rays(prim, sample_id=rank))
        s_parts.append(fx.shadow_rays(prim))
    diffuse = np.concatenate(d_parts); shadow = np.concatenate(s_parts)
    del d_parts, s_parts
    nd, ns = len(diffuse), len(shadow)
    h_d = torch.from_numpy(diffuse.view(np.uint8).reshape(nd, 80)).pin_memory()
    h_s = torch.from_numpy(shadow.view(np.uint8).reshape(ns, 48)).pin_memory()
    p_d, p_s = h_d.cuda(), h_s.cuda()                                       # pristine device copies
    w_d, w_s = torch.empty_like(p_d), torch.empty_like(p_s)                 # working copies (traced in place)

    def step_device():
        lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80)
        lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)

    # ---- counters (instrumented kernels, outside the timed region): the roofline numerator ----
    w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
    c_close = lib.intersect_counted(sc, w_d.data_ptr(), nd, 80)
    c_occ = lib.intersect_counted(sc, w_s.data_ptr(), ns, 48, occluded=True)
    hits = int(c_close["hits"])

    # ---- timed region: device-resident streams, CUDA events on the launching stream ----
    for _ in range(args.warmup):
        w_d.copy_(p_d); w_s.copy_(p_s); step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = lib.lib.rtcxGetLaunchCount()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(args.steps)]
    for k in range(args.steps):
        w_d.copy_(p_d); w_s.copy_(p_s)                                      # fresh inputs (also flushes L2: 2.1 GB >> 126 MB); not timed
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

    # ---- end to end through the C ABI with pinned host buffers (H2D + kernels + D2H timed) ----
    e2e_steps = max(1, min(args.steps, 5))
    hw_d, hw_s = torch.empty_like(h_d).pin_memory(), torch.empty_like(h_s).pin_memory()
    t_e2e = 0.0
    xfer0 = (0, 0)
    for k in range(1 + e2e_steps):
        if k == 1:
            xfer0 = lib.transfer_bytes(dev)                                # bytes the library itself copies over PCIe (counted at its cudaMemcpy calls)
        hw_d.copy_(h_d); hw_s.copy_(h_s)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        lib.intersect_ptr(sc, hw_d.data_ptr(), nd, 80)                      # host pointer: staged by the library
        lib.occluded_ptr(sc, hw_s.data_ptr(), ns, 48)
        dt = time.perf_counter() - t0
        if k > 0:
            t_e2e += dt
    t_e2e /= e2e_steps
    xfer1 = lib.transfer_bytes(dev)
    h2d_step, d2h_step = (xfer1[0] - xfer0[0]) // e2e_steps, (xfer1[1] - xfer0[1]) // e2e_steps
    clocks = sampler.stop()
    same = bool(np.array_equal(hw_d.numpy(), w_d.cpu().numpy()))             # host path == device path, bit for bit

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([ms_step, t_e2e * 1e3, t_close, t_occ], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, t_close, t_occ = [float(x) for x in t.tolist()]
        cnt = torch.tensor([nd + ns, launches, h2d_step, d2h_step], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        total_rays, launches, h2d_step, d2h_step = int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3])
    else:
        e2e_ms, total_rays = t_e2e * 1e3, nd + ns

    if rank == 0:
        peaks = measured_peaks()
        peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        node_b, tri_b = 80, 48                                             # bytes of a node / triangle record traversal reads
        alg_close = c_close["nodes"] * node_b + c_close["tris"] * tri_b + nd * 48 + hits * 36
        ach = alg_close / (t_close * 1e-3) / 1e9
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_rays / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, ntris), "rays_per_gpu_per_step": nd + ns,
                       "l2_policy": "inputs (2.1 GB of rays per step) exceed L2; streams re-copied from pristine buffers between timed steps",
                       "parallelism": f"ray-sharded x{world}, BVH replica per GPU"},
            "closest_mrays_per_s": world * nd / (t_close * 1e-3) / 1e6, "occluded_mrays_per_s": world * ns / (t_occ * 1e-3) / 1e6,
            "build": build, "build_mtris_per_s": (ntris / (build["msTotal"] * 1e-3) / 1e6) if build else None,
            "bvh_broadcast_ms": bcast_ms,
            "traversal_per_ray": {"closest_nodes": c_close["nodes"] / nd, "closest_tris": c_close["tris"] / nd,
                                  "occluded_nodes": c_occ["nodes"] / max(c_occ["rays"], 1), "occluded_tris": c_occ["tris"] / max(c_occ["rays"], 1),
                                  "hit_fraction": hits / nd},
            "roofline": {"bound": "hbm", "kernel": "k_trace<closest>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "bytes_per_ray": alg_close / nd, "achieved_if_nodes_count_128B": (alg_close + c_close["nodes"] * 48) / (t_close * 1e-3) / 1e9,
                         "traffic": ncu.get("dram_bytes_per_launch_closest") if args.workload == "c2" else None},   # the ncu capture is of the configs[1] streams
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "host_record_bytes_per_step": world * (nd * 80 + ns * 48),
                    "path": "rtcIntersect1M + rtcOccluded1M on page-locked host streams: H2D of the ray records, kernels, compact "
                            "hit-list D2H, scatter into the caller's records by the library's host threads (all inside the timed region)",
                    "host_equals_device_result": same},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                S = reference_setup(fx, rt, args.workload, rows=(1792, 2304))
                if S is not None:
                    reference_step(S)
                    tt, nn = 0.0, 0
                    for _ in range(3):
                        t, n = reference_step(S)
                        tt += t; nn += n
                    line["cpu_baseline"] = {"value": nn / tt / 1e6, "unit": "Mrays/s", "cores": S["cores"], "kind": "reference",
                                            "sample": f"3 x ({len(S['diffuse'])} diffuse + {len(S['shadow'])} shadow rays), rows 1792-2303 of the frame, 4096-ray chunks per thread",
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
