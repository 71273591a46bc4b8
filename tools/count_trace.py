import importlib, sys
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore
lib = rt.RTCore(); dev = lib.new_device(sys.argv[1] if len(sys.argv) > 1 else "")
sc, keep = lib.build_scene(dev, fx.scene_c2(1.0))
prim = fx.primary_rays(4096, 4096, rows=(1792, 2304), **fx.C2_CAMERA)
lib.intersect(sc, prim, coherent=True)
d = fx.diffuse_rays(prim); s = fx.shadow_rays(prim)
c = lib.intersect_counted(sc, d); n = c["rays"]
print("closest", {k: round(v / n, 3) for k, v in c.items()}, "nodes/hit-ray", c["hitNodes"] / max(c["hits"], 1), "nodes/miss-ray", (c["nodes"] - c["hitNodes"]) / max(n - c["hits"], 1))
c = lib.intersect_counted(sc, s, occluded=True); n = c["rays"]
print("occluded", {k: round(v / n, 3) for k, v in c.items()})
