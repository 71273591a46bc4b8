"""Per-ray work of the reference ALGORITHM (oracle restatement: binned-SAH BVH8 of Triangle4 blocks, single-ray
traversal, nearest-first order, cull on pop) on the bench streams -- the target our node/triangle counts are compared with.
Measurement tooling; runs on the CPU."""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx = pkg.fixtures
from oracle.rq_oracle import Oracle
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 64
o = Oracle()
meshes = fx.scene_c3(1.0) if w == "c3" else fx.scene_c2(1.0)
t0 = time.time(); h = o.build(meshes); print(w, "oracle build %.1fs" % (time.time() - t0), o.stats(h), "sah", o.sah(h), flush=True)
prim = fx.primary_rays(4096, 4096, rows=(1792, 1792 + rows), **fx.C2_CAMERA)
o.trace_counters(); o.intersect(h, prim); c = o.trace_counters()
print("primary per ray:", {k: round(v / c["rays"], 3) for k, v in c.items()})
d = fx.diffuse_rays(prim); s = fx.shadow_rays(prim)
o.intersect(h, d); c = o.trace_counters()
print("diffuse per ray:", {k: round(v / c["rays"], 3) for k, v in c.items()}, "hit fraction", float((d["geomID"] != 0xFFFFFFFF).mean()))
o.occluded(h, s); c = o.trace_counters()
print("shadow per ray:", {k: round(v / max(c["rays"], 1), 3) for k, v in c.items()})
