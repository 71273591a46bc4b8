"""Generates tests/golden/*.npz from the REAL reference library (oracle/_ref/libembree3_ref.so, built by
oracle/build_ref.py from /root/reference).  Run here (the container that has the reference); the
.npz files are committed so that tests on machines without the reference still pin the oracle and
the product.  Stored per case: meshes, input rays, the reference's rtcIntersect1M output, a shadow
stream derived from those hits with the reference's rtcOccluded1M output, rtcOccluded1M applied to
the input rays themselves, and the reference's own SAH (BENCHMARK_BUILD line) and scene bounds.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402

rt, fx = cases.rt, cases.fx
REF = os.path.join(cases.ROOT, "oracle", "_ref", "libembree3_ref.so")


def capture_stdout(fn):
    sys.stdout.flush()
    with tempfile.TemporaryFile() as tmp:
        old = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            r = fn()
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(old, 1)
            os.close(old)
        tmp.seek(0)
        return r, tmp.read().decode()


def main():
    ref = rt.RTCore(REF)
    dev = ref.new_device("benchmark=1")
    for name, make in cases.CASES.items():
        c = make()
        (sc, keep), out = capture_stdout(lambda: ref.build_scene(dev, c["meshes"], c["flags"]))
        m = re.search(r"BENCHMARK_BUILD\s+\S+\s+\S+\s+(\S+)\s+(\S+)", out)
        sah = float(m.group(1)) if m else float("nan")
        b = rt.Bounds()
        ref.lib.rtcGetSceneBounds(sc, C.byref(b))
        bounds = np.array([b.lower_x, b.lower_y, b.lower_z, b.upper_x, b.upper_y, b.upper_z], dtype=np.float32)
        rays_in = c["rays"]
        closest = rays_in.copy()
        ref.intersect(sc, closest)
        shadow_in = fx.shadow_rays(closest, light=(3.0, 6.0, 2.0)) if (closest["geomID"] != 0xFFFFFFFF).any() else fx.to_ray(rays_in[:0])
        shadow_out = shadow_in.copy()
        if len(shadow_out):
            ref.occluded(sc, shadow_out)
        occl_self = fx.to_ray(rays_in)
        if "groups" in c:                      # one stream per category (see cases.case_edge_rays)
            for gid in np.unique(c["groups"]):
                part = occl_self[c["groups"] == gid].copy()
                ref.occluded(sc, part)
                occl_self[c["groups"] == gid] = part
        else:
            ref.occluded(sc, occl_self)
        d = dict(num_meshes=len(c["meshes"]), flags=c["flags"], rays_in=rays_in.view(np.uint8), closest_out=closest.view(np.uint8),
                 shadow_in=shadow_in.view(np.uint8), shadow_out=shadow_out.view(np.uint8), occl_self_out=occl_self.view(np.uint8),
                 sah_ref=sah, bounds_ref=bounds)
        for i, (v, t) in enumerate(c["meshes"]):
            d[f"v{i}"], d[f"t{i}"] = np.asarray(v, np.float32), np.asarray(t, np.uint32)
        if "expect_uv" in c:
            d["expect_u"], d["expect_v"] = c["expect_uv"]
        if "expect_prim" in c:
            d["expect_prim"] = c["expect_prim"]
        if "groups" in c:
            d["groups"] = c["groups"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        hits = int((closest["geomID"] != 0xFFFFFFFF).sum())
        print(f"{name}: tris={fx.num_tris(c['meshes'])} rays={len(rays_in)} hits={hits} shadow={len(shadow_in)} "
              f"occluded={int(np.isneginf(shadow_out['tfar']).sum())} sah_ref={sah}")
        ref.lib.rtcReleaseScene(sc)


if __name__ == "__main__":
    main()
