"""Profiling harness: prepares the configs[1] (or c3) streams on the device, then brackets exactly one
closest-hit launch and one occlusion launch with cudaProfilerStart/Stop (use ncu --profile-from-start off)."""
import argparse, importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--bands", type=int, default=8)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--cfg", default="")
ap.add_argument("--counters", action="store_true")
ap.add_argument("--sort", default="none", help="host-side reordering experiment: none|octant|chunkdir:<chunk>:<bins>|global:<obits>:<bins>")
ap.add_argument("--lib", default=None, help="path of an experiment build (csrc/Makefile variant)")
ap.add_argument("--meta", default=None, help="write {kernel_source_hash, git, rays_per_launch} here (read by tools/summarize_profiles.py)")
args = ap.parse_args()
lib = rt.RTCore(args.lib) if args.lib else rt.RTCore()
dev = lib.new_device("async=1," + args.cfg)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
lib.lib.rtcxSetDeviceStream(dev, st.cuda_stream)
meshes = fx.scene_c3(1.0) if args.workload == "c3" else fx.scene_c2(1.0)
sc, keep = lib.build_scene(dev, meshes)
print("build", lib.build_stats(sc), flush=True)
d_parts, s_parts = [], []
for b in range(args.bands):
    prim = fx.primary_rays(4096, 4096, rows=(b * 512, b * 512 + 512), **fx.C2_CAMERA)
    lib.intersect(sc, prim, coherent=True)
    d_parts.append(fx.diffuse_rays(prim)); s_parts.append(fx.shadow_rays(prim))
diffuse, shadow = np.concatenate(d_parts), np.concatenate(s_parts)


def dir_bin(r, nb):
    """octahedral map of the direction to an nb x nb grid"""
    d = np.stack([r["dir_x"], r["dir_y"], r["dir_z"]], 1).astype(np.float64)
    d /= np.maximum(np.abs(d).sum(1, keepdims=True), 1e-30)
    u, v = d[:, 0].copy(), d[:, 1].copy()
    neg = d[:, 2] < 0
    uu = (1 - np.abs(v)) * np.where(u >= 0, 1, -1); vv = (1 - np.abs(u)) * np.where(v >= 0, 1, -1)
    u = np.where(neg, uu, u); v = np.where(neg, vv, v)
    iu = np.clip(((u * 0.5 + 0.5) * nb).astype(np.int64), 0, nb - 1); iv = np.clip(((v * 0.5 + 0.5) * nb).astype(np.int64), 0, nb - 1)
    return iu * nb + iv


def reorder(r, mode):
    if mode == "none":
        return r
    n = len(r)
    if mode == "octant":
        key = (r["dir_x"] < 0).astype(np.int64) | ((r["dir_y"] < 0).astype(np.int64) << 1) | ((r["dir_z"] < 0).astype(np.int64) << 2)
    elif mode.startswith("chunkdir"):
        _, chunk, nb = mode.split(":")
        key = (np.arange(n) // int(chunk)) * 100000 + dir_bin(r, int(nb))
    elif mode.startswith("global"):
        _, ob, nb = mode.split(":")
        ob = int(ob)
        o = np.stack([r["org_x"], r["org_y"], r["org_z"]], 1).astype(np.float64)
        lo, hi = o.min(0), o.max(0)
        q = np.clip(((o - lo) / np.maximum(hi - lo, 1e-30) * (1 << ob)).astype(np.int64), 0, (1 << ob) - 1)
        m = np.zeros(n, dtype=np.int64)
        for b in range(ob):
            for a in range(3):
                m |= ((q[:, a] >> b) & 1) << (3 * b + a)
        key = m * 100000 + dir_bin(r, int(nb))
    return r[np.argsort(key, kind="stable")]


diffuse, shadow = reorder(diffuse, args.sort), reorder(shadow, args.sort)
nd, ns = len(diffuse), len(shadow)
if args.meta:
    import json, subprocess
    sys.path.insert(0, ".")
    import bench
    try:
        git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except Exception:
        git = None
    json.dump({"kernel_source_hash": bench.kernel_source_hash(), "git": git, "rays_per_launch": nd, "workload": args.workload}, open(args.meta, "w"))
p_d = torch.from_numpy(diffuse.view(np.uint8).reshape(nd, 80)).cuda()
p_s = torch.from_numpy(shadow.view(np.uint8).reshape(ns, 48)).cuda()
w_d, w_s = p_d.clone(), p_s.clone()
lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80); lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)
torch.cuda.synchronize()
if args.counters:
    w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
    c = lib.intersect_counted(sc, w_d.data_ptr(), nd, 80)
    print("closest per ray:", {k: round(v / max(c["rays"], 1), 3) for k, v in c.items() if k != "stackMax"}, "stackMax", c["stackMax"], flush=True)
    c = lib.intersect_counted(sc, w_s.data_ptr(), ns, 48, occluded=True)
    print("occluded per ray:", {k: round(v / max(c["rays"], 1), 3) for k, v in c.items() if k != "stackMax"}, "stackMax", c["stackMax"], flush=True)
for rep in range(args.reps):
    w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
    e = [torch.cuda.Event(True) for _ in range(3)]
    torch.cuda.profiler.start()
    e[0].record(); lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80)
    e[1].record(); lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)
    e[2].record(); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("closest %.3f ms %.1f Mrays/s | occluded %.3f ms %.1f Mrays/s" % (e[0].elapsed_time(e[1]), nd / e[0].elapsed_time(e[1]) / 1e3,
          e[1].elapsed_time(e[2]), ns / e[1].elapsed_time(e[2]) / 1e3), flush=True)
