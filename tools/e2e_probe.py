"""Focused end-to-end probe: the configs[1] streams through rtcIntersect1M / rtcOccluded1M on page-locked host buffers,
a few repetitions per device configuration (much cheaper than bench.py when only `e2e` is of interest).
usage: python tools/e2e_probe.py "<cfg>" "<cfg>" ...      ("-" = defaults; add verbose=2 for the library's own breakdown)"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures

lib = rt.RTCore()
meshes = fx.scene_c2(1.0)
dev0 = lib.new_device("")
sc0, keep0 = lib.build_scene(dev0, meshes)
d_parts, s_parts = [], []
for b in range(8):
    prim = fx.primary_rays(4096, 4096, rows=(b * 512, (b + 1) * 512), **fx.C2_CAMERA)
    lib.intersect(sc0, prim, coherent=True)
    d_parts.append(fx.diffuse_rays(prim, sample_id=0)); s_parts.append(fx.shadow_rays(prim))
diffuse, shadow = np.concatenate(d_parts), np.concatenate(s_parts)
nd, ns = len(diffuse), len(shadow)
h_d = torch.from_numpy(diffuse.view(np.uint8).reshape(nd, 80)).pin_memory()
h_s = torch.from_numpy(shadow.view(np.uint8).reshape(ns, 48)).pin_memory()
w_d, w_s = torch.empty_like(h_d).pin_memory(), torch.empty_like(h_s).pin_memory()
ref_d = ref_s = None
for cfg in sys.argv[1:] or ["-"]:
    c = "" if cfg == "-" else cfg
    dev = lib.new_device(c)
    sc, keep = lib.build_scene(dev, meshes)
    ts = []
    for k in range(4):
        w_d.copy_(h_d); w_s.copy_(h_s)
        torch.cuda.synchronize()
        time.sleep(float(os.environ.get("PROBE_SETTLE", "0")))          # optional settle time after the host-side reset copy (PROBE_SETTLE seconds): measured, no consistent effect
        t0 = time.perf_counter()
        lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80)
        t1 = time.perf_counter()
        lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48)
        t2 = time.perf_counter()
        ts.append((t2 - t0, t1 - t0, t2 - t1))
    if ref_d is None:
        ref_d, ref_s = w_d.numpy().copy(), w_s.numpy().copy()
    same = bool(np.array_equal(ref_d, w_d.numpy()) and np.array_equal(ref_s, w_s.numpy()))
    best = min(ts[1:])
    x = lib.transfer_bytes(dev)
    print(f"{cfg:44s} e2e {(nd + ns) / best[0] / 1e6:7.1f} Mrays/s  closest {best[1] * 1e3:6.2f} ms  occluded {best[2] * 1e3:6.2f} ms  "
          f"h2d {x[0] / 4 / 1e9:.3f} GB d2h {x[1] / 4 / 1e9:.3f} GB per step  same_as_first={same}  err={lib.lib.rtcGetDeviceError(dev)}", flush=True)
    lib.lib.rtcReleaseScene(sc); lib.lib.rtcReleaseDevice(dev)
