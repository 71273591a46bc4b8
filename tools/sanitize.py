"""Small end-to-end exercise for compute-sanitizer (memcheck / racecheck / initcheck): build (both front ends), refit, flat and
instanced traversal (closest + occluded, device-resident and host-staged compact path), image export / import.
usage: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures
import instancing  # noqa: E402

lib = rt.RTCore()
for cfg in ("gpu_builder=sah,treelet=256", "gpu_builder=sah", "gpu_builder=ploc", "gpu_builder=lbvh,chunk_rays=20000"):
    dev = lib.new_device(cfg)
    meshes = [fx.displaced_plane(40, extent=3.0), fx.triangle_sphere((0, 1, 0), 0.7, 16)]
    sc = lib.lib.rtcNewScene(dev)
    keep, geoms = [], []
    for v, t in meshes:
        _, g = lib.add_mesh(dev, sc, v, t, keep)
        lib.lib.rtcSetGeometryBuildQuality(g, rt.RTC_BUILD_QUALITY_REFIT); lib.lib.rtcCommitGeometry(g); geoms.append(g)
    lib.lib.rtcCommitScene(sc)
    rays = np.tile(fx.incoherent_rays(20000, org=(0.1, 2.0, 0.2), seed=3), 4)      # 80000 rays: the compact staged path engages
    a = rays.copy(); lib.intersect(sc, a)
    s = fx.shadow_rays(a); lib.occluded(sc, s)
    d = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), 80).copy()).cuda()
    lib.intersect_ptr(sc, d.data_ptr(), len(rays))
    assert np.array_equal(d.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE), a)
    keep[0][:meshes[0][0].size] += 0.05
    lib.lib.rtcUpdateGeometryBuffer(geoms[0], rt.RTC_BUFFER_TYPE_VERTEX, 0); lib.lib.rtcCommitGeometry(geoms[0]); lib.lib.rtcCommitScene(sc)
    assert lib.build_stats(sc)["refitCount"] == 1
    b = rays.copy(); lib.intersect(sc, b)
    n = C.c_size_t(); p = lib.lib.rtcxGetSceneImage(sc, C.byref(n))
    sc2 = lib.lib.rtcNewScene(dev); lib.lib.rtcxSetSceneImage(sc2, p, n.value)
    c = rays.copy(); lib.intersect(sc2, c)
    assert np.array_equal(b, c)
    cs = instancing.CASES["inst_forest"]()
    top, objs, k2 = lib.build_instanced(dev, cs["objects"], cs["base"], cs["instances"], 0)
    r = np.tile(cs["rays"], 6); lib.intersect(top, r); sh = fx.shadow_rays(r); lib.occluded(top, sh)
    assert lib.lib.rtcGetDeviceError(dev) == 0
    print(cfg, "hits", int((a["geomID"] != 0xFFFFFFFF).sum()), int((r["geomID"] != 0xFFFFFFFF).sum()), flush=True)
    for x in [top] + objs + [sc, sc2]:
        lib.lib.rtcReleaseScene(x)
    lib.lib.rtcReleaseDevice(dev)
print("sanitize run ok")
