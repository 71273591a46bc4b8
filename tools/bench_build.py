"""BASELINE.json configs[4]: GPU BVH build of 10 M - 50 M triangles (structured: the configs[2] scene scaled by tessellation;
unstructured: random small triangles in the unit cube, seed 7) with both front ends, the reference's binned-SAH BVH8 builder on
the host cores beside it, and -- as "traversal quality" -- the SAH figure plus a 4 M-ray incoherent probe stream on each tree.
usage: python tools/bench_build.py [--sizes 10,20,50] [--no-reference] [--ref-max 20]"""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures


def probe_rays(n, lo, hi, seed=5):
    rs = fx.RandomSampler(np.arange(n), seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    o = np.stack([lo[k] + (hi[k] - lo[k]) * rs.get_float() for k in range(3)], 1).astype(np.float32)
    d = np.stack([rs.get_float() * 2 - 1 for _ in range(3)], 1).astype(np.float32)
    return fx._set(rt.new_rays(n), o, d, 0.0, np.inf)


def ours(lib, meshes, cfg, rays):
    import torch
    dev = lib.new_device(cfg)
    sc, keep = lib.build_scene(dev, meshes)                     # first commit: pool growth included
    walls = []
    for _ in range(2):
        for g in range(len(meshes)):
            lib.lib.rtcCommitGeometry(lib.lib.rtcGetGeometry(sc, g))
        t0 = time.perf_counter(); lib.lib.rtcCommitScene(sc); walls.append((time.perf_counter() - t0) * 1e3)
    st = lib.build_stats(sc)
    assert lib.lib.rtcGetDeviceError(dev) == 0
    d = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), 80).copy()).cuda()
    w = d.clone()
    best = 1e9
    for _ in range(3):
        w.copy_(d); torch.cuda.synchronize()
        t0 = time.perf_counter(); lib.intersect_ptr(sc, w.data_ptr(), len(rays)); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out = w.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE)
    res = {"device_ms": st["msTotal"], "mtris_per_s_device": st["numPrimsValid"] / st["msTotal"] / 1e3, "commit_wall_ms": min(walls),
           "mtris_per_s_wall": st["numPrimsValid"] / min(walls) / 1e3, "sah": st["sah"], "sahExact": st["sahExact"], "nodes": st["numNodes"],
           "sahInner": st["sahInner"], "sahLeafTris": st["sahLeafTris"], "sah_blocks_of_4_equivalent": st["sahInner"] + st["sahLeafTris"] / 4.0,
           "depth": st["depth"], "image_gb": st["bytes"] / 1e9, "phases_ms": {k: st[k] for k in ("msPrims", "msSort", "msHierarchy", "msRefit", "msEmit")},
           "probe_mrays_per_s": len(rays) / best / 1e6, "probe_hits": int((out["geomID"] != 0xFFFFFFFF).sum())}
    lib.lib.rtcReleaseScene(sc); lib.lib.rtcReleaseDevice(dev)
    torch.cuda.empty_cache()
    return res, out


def capture_stdout(fn):
    sys.stdout.flush()
    with tempfile.TemporaryFile() as tmp:
        old = os.dup(1); os.dup2(tmp.fileno(), 1)
        try:
            r = fn()
        finally:
            C.CDLL(None).fflush(None); os.dup2(old, 1); os.close(old)
        tmp.seek(0)
        return r, tmp.read().decode()


def reference_build(ref, meshes, rays):
    dev = ref.new_device("benchmark=1")
    t0 = time.perf_counter()
    (sc, keep), out = capture_stdout(lambda: ref.build_scene(dev, meshes))
    wall = (time.perf_counter() - t0) * 1e3
    m = re.search(r"BENCHMARK_BUILD\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)", out)
    r = rays[:1 << 20].copy()
    import threading
    nth = os.cpu_count() or 1
    parts = np.array_split(np.arange(len(r)), nth)

    def work(idx):
        p = r[idx[0]:idx[-1] + 1]
        for c0 in range(0, len(p), 4096):
            ref.intersect(sc, p[c0:c0 + 4096])
    th = [threading.Thread(target=work, args=(p,)) for p in parts if len(p)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    res = {"build_s_reported": float(m.group(1)) if m else None, "mtris_per_s": float(m.group(2)) / 1e6 if m else None,
           "sah": float(m.group(3)) if m else None, "bytes": int(float(m.group(4))) if m else None, "build_wall_ms_incl_attach": wall,
           "probe_mrays_per_s": len(r) / dt / 1e6, "threads": nth}
    ref.lib.rtcReleaseScene(sc); ref.lib.rtcReleaseDevice(dev)
    return res, r


LAUNCHES = [0]
BUILDERS = (("ours", ""), ("ploc", "gpu_builder=ploc"), ("lbvh", "gpu_builder=lbvh"))


def run(sizes=(10.0, 20.0, 50.0), kinds=("scene", "soup"), reference=True, ref_max=20.0, builders=BUILDERS, emit=None):
    """One row per (kind, size): every builder front end ('ours' = the device default), the reference's builder beside it."""
    parity = importlib.import_module("embree-aarch64_b200.parity")
    lib = rt.RTCore()
    from oracle.rq_oracle import REF_LIB
    ref = rt.RTCore(REF_LIB) if (reference and os.path.exists(REF_LIB)) else None
    rows = []
    for kind in kinds:
        for m in sizes:
            if kind == "scene":
                meshes = fx.scene_c3(float(np.sqrt(m / 10.0)))
                lo, hi = (-9.0, 0.5, -9.0), (9.0, 4.0, 9.0)
            else:
                meshes = fx.random_soup(int(m * 1e6))
                lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
            n = fx.num_tris(meshes)
            rays = probe_rays(1 << 22, lo, hi)
            line = {"kind": kind, "triangles": n}
            outs = {}
            for name, cfg in builders:
                line[name], outs[name] = ours(lib, meshes, cfg, rays)
            if ref is not None and m <= ref_max:
                line["reference"], rr = reference_build(ref, meshes, rays)
                c = parity.compare_closest(outs[builders[0][0]][:len(rr)], rr)
                line["parity_vs_reference_1M_rays"] = {k: c[k] for k in ("pass", "agreement", "hitmiss_disagree", "id_disagree_unexplained", "max_t_rel", "max_uv_abs")}
            rows.append(line)
            if emit:
                emit(line)
            del meshes, rays, outs
    LAUNCHES[0] = int(lib.lib.rtcxGetLaunchCount())
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10,20,50")
    ap.add_argument("--kinds", default="scene,soup")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--ref-max", type=float, default=20)
    a = ap.parse_args()
    run([float(x) for x in a.sizes.split(",")], a.kinds.split(","), not a.no_reference, a.ref_max, emit=lambda l: print(json.dumps(l), flush=True))


if __name__ == "__main__":
    main()
