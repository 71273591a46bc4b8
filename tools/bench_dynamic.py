"""Dynamic-scene commit cost: full rebuild vs refit of the configs[1] scene after a
vertex update, ours on the GPU and -- beside it -- the reference library on the host cores
(RTC_SCENE_FLAG_DYNAMIC + RTC_BUILD_QUALITY_REFIT, the setting its refit path needs).
usage: python tools/bench_dynamic.py [--scale 1.0] [--reps 5] [--no-reference]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures


def wave(v0, phase):
    v = v0.copy()
    v[:, 1] += (0.2 * np.sin(1.3 * v0[:, 0] + phase) * np.cos(0.9 * v0[:, 2] - 0.5 * phase)).astype(np.float32)
    return v


def run(lib, dev, meshes, reps, quality, flags, device_buffers=False):
    L = lib.lib
    sc = L.rtcNewScene(dev)
    if flags:
        L.rtcSetSceneFlags(sc, flags)
    keep, geoms, bufs = [], [], []
    if device_buffers:
        import torch
        for v, t in meshes:
            vpad = np.zeros(v.size + 4, dtype=np.float32)
            vpad[:v.size] = np.asarray(v, dtype=np.float32).ravel()
            dv = torch.from_numpy(vpad).cuda()
            dt = torch.from_numpy(np.ascontiguousarray(t, dtype=np.uint32).view(np.int32)).cuda()
            g = L.rtcNewGeometry(dev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
            L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, dv.data_ptr(), 0, 12, len(v))
            L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, dt.data_ptr(), 0, 12, len(t))
            L.rtcSetGeometryBuildQuality(g, quality)
            L.rtcCommitGeometry(g)
            L.rtcAttachGeometry(sc, g)
            keep += [dv, dt]
            geoms.append(g)
            bufs.append(dv)
    else:
        for v, t in meshes:
            _, g = lib.add_mesh(dev, sc, v, t, keep)
            L.rtcSetGeometryBuildQuality(g, quality)
            L.rtcCommitGeometry(g)
            geoms.append(g)
        bufs = [keep[0], keep[2]] if len(meshes) == 2 else [keep[2 * i] for i in range(len(meshes))]
    t0 = time.perf_counter()
    L.rtcCommitScene(sc)
    first = (time.perf_counter() - t0) * 1e3
    wall, devms, refits = [], [], 0
    for r in range(reps):
        phase = 0.3 * (r + 1)
        for (v, t), b in zip(meshes, bufs):
            nv = wave(np.asarray(v, dtype=np.float32), phase)
            if device_buffers:
                import torch
                b[:nv.size].copy_(torch.from_numpy(nv.ravel()))
                torch.cuda.synchronize()
            else:
                b[:nv.size] = nv.ravel()
        for g in geoms:
            L.rtcUpdateGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0)
            L.rtcCommitGeometry(g)
        t0 = time.perf_counter()
        L.rtcCommitScene(sc)
        wall.append((time.perf_counter() - t0) * 1e3)
        if lib.has_ext:
            st = lib.build_stats(sc)
            devms.append(st["msTotal"])
            refits = st["refitCount"]
    # sanity: the updated scene still answers rays
    rays = fx.primary_rays(64, 64, org=(0, 8, 0), look=(0, -1, 0), up=(0, 0, 1))
    lib.intersect(sc, rays)
    hits = int((rays["geomID"] != 0xFFFFFFFF).sum())
    err = L.rtcGetDeviceError(dev)
    for g in geoms:
        L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)
    ntris = fx.num_tris(meshes)
    w = float(np.median(wall))
    out = {"first_commit_ms": first, "commit_wall_ms_median": w, "commit_wall_ms_best": float(min(wall)), "mtris_per_s_wall": ntris / w / 1e3,
           "refitCount": refits, "hits": hits, "error": err}
    if devms:
        d = float(np.median(devms))
        out.update({"device_ms_median": d, "mtris_per_s_device": ntris / d / 1e3})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--cfg", default="", help="device options of our library (e.g. stage_geometry=0)")
    a = ap.parse_args()
    meshes = fx.scene_c2(a.scale)
    ntris = fx.num_tris(meshes)
    lib = rt.RTCore()
    dev = lib.new_device(a.cfg)
    res = {"workload": f"configs[1] scene, {ntris} triangles, wave deformation of every vertex per commit", "reps": a.reps}
    res["ours_rebuild_host_buffers"] = run(lib, dev, meshes, a.reps, rt.RTC_BUILD_QUALITY_MEDIUM, 0)
    res["ours_refit_host_buffers"] = run(lib, dev, meshes, a.reps, rt.RTC_BUILD_QUALITY_REFIT, 0)
    res["ours_rebuild_device_buffers"] = run(lib, dev, meshes, a.reps, rt.RTC_BUILD_QUALITY_MEDIUM, 0, device_buffers=True)
    res["ours_refit_device_buffers"] = run(lib, dev, meshes, a.reps, rt.RTC_BUILD_QUALITY_REFIT, 0, device_buffers=True)
    lib.lib.rtcReleaseDevice(dev)
    if not a.no_reference:
        from oracle.rq_oracle import REF_LIB
        if os.path.exists(REF_LIB):
            ref = rt.RTCore(REF_LIB)
            rdev = ref.new_device("")
            res["reference_threads"] = os.cpu_count()
            res["reference_rebuild_static"] = run(ref, rdev, meshes, max(2, a.reps // 2), rt.RTC_BUILD_QUALITY_MEDIUM, 0)
            res["reference_refit_dynamic"] = run(ref, rdev, meshes, max(2, a.reps // 2), rt.RTC_BUILD_QUALITY_REFIT, rt.RTC_SCENE_FLAG_DYNAMIC)
            ref.lib.rtcReleaseDevice(rdev)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
