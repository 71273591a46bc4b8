"""Builds one scene and prints the build statistics: the target of ncu captures of the builder kernels, and the quick same-box
A/B of two library builds on build time (phases in ms) + answers.
usage: python tools/profile_build.py [--workload c3|c2] [--lib <variant .so>] [--cfg <device options>] [--commits 3]"""
import argparse, importlib, json, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--lib", default=None)
ap.add_argument("--cfg", default="")
ap.add_argument("--commits", type=int, default=3)
args = ap.parse_args()
lib = rt.RTCore(args.lib) if args.lib else rt.RTCore()
dev = lib.new_device(args.cfg)
meshes = fx.scene_c3(1.0) if args.workload == "c3" else fx.scene_c2(1.0)
sc, keep = lib.build_scene(dev, meshes)
for k in range(args.commits):
    if k:
        for g in range(len(meshes)):
            lib.lib.rtcCommitGeometry(lib.lib.rtcGetGeometry(sc, g))
        t0 = time.perf_counter(); lib.lib.rtcCommitScene(sc); wall = (time.perf_counter() - t0) * 1e3
    else:
        wall = None
    st = lib.build_stats(sc)
    print(json.dumps({"commit": k, "wall_ms": wall, **{x: st[x] for x in ("msTotal", "msPrims", "msSort", "msHierarchy", "msRefit", "msEmit", "numNodes", "depth", "sah", "numTreelets")}}), flush=True)
# answers: a primary pass + its diffuse bounce, checksummed (compare between library builds)
prim = fx.primary_rays(1024, 1024, **fx.C2_CAMERA)
lib.intersect(sc, prim, coherent=True)
d = fx.diffuse_rays(prim)
lib.intersect(sc, d)
import hashlib
print(json.dumps({"hits_primary": int((prim["geomID"] != 0xFFFFFFFF).sum()), "hits_diffuse": int((d["geomID"] != 0xFFFFFFFF).sum()),
                  "sha_primary": hashlib.sha1(prim.tobytes()).hexdigest()[:12], "sha_diffuse": hashlib.sha1(d.tobytes()).hexdigest()[:12],
                  "error": lib.lib.rtcGetDeviceError(dev)}))
