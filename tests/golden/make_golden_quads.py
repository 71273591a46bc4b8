"""Golden vectors for quad meshes, produced by the REAL reference library (oracle/_ref/libembree3_ref.so built with
EMBREE_GEOMETRY_QUAD, see oracle/build_ref.py).
    python tests/golden/make_golden_quads.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import quads  # noqa: E402

rt, fx = cases.rt, cases.fx
REF = os.path.join(cases.ROOT, "oracle", "_ref", "libembree3_ref.so")


def main():
    ref = rt.RTCore(REF)
    dev = ref.new_device("")
    for name, make in quads.CASES.items():
        c = make()
        sc, keep = ref.build_scene(dev, c["meshes"], c["flags"])
        assert ref.lib.rtcGetDeviceError(dev) == 0
        b = rt.Bounds()
        ref.lib.rtcGetSceneBounds(sc, C.byref(b))
        bounds = np.array([b.lower_x, b.lower_y, b.lower_z, b.upper_x, b.upper_y, b.upper_z], dtype=np.float32)
        closest = c["rays"].copy()
        ref.intersect(sc, closest)
        shadow_in = fx.shadow_rays(closest, light=(3.0, 8.0, 2.0))
        shadow_out = shadow_in.copy()
        ref.occluded(sc, shadow_out)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rays_in=c["rays"].view(np.uint8), closest_out=closest.view(np.uint8),
                            shadow_in=shadow_in.view(np.uint8), shadow_out=shadow_out.view(np.uint8), bounds_ref=bounds)
        hit = closest["geomID"] != 0xFFFFFFFF
        print(f"{name}: rays={len(closest)} hits={int(hit.sum())} on quads={int((closest['geomID'][hit] != 1).sum())} "
              f"shadow={len(shadow_in)} occluded={int(np.isneginf(shadow_out['tfar']).sum())} bounds={bounds}")
        ref.lib.rtcReleaseScene(sc)


if __name__ == "__main__":
    main()
