"""Quad-mesh parity cases (RTC_GEOMETRY_TYPE_QUAD as two triangles), shared by the golden generator
(tests/golden/make_golden_quads.py), the oracle tests and the GPU tests.  Reference behaviour: a quad (v0,v1,v2,v3) is
intersected as the triangles (v0,v1,v3) and (v2,v3,v1), the second reporting u = 1-u, v = 1-v; primID = quad index
(kernels/geometry/quad_intersector_moeller.h:122-144)."""
import numpy as np

import cases

fx, rt = cases.fx, cases.rt


def bumpy_quads(n, extent, seed=3):
    """n x n quads over [-extent, extent]^2, every vertex displaced in y: the quads are NOT planar."""
    v, q = fx.quad_plane((-extent, 0, -extent), (2 * extent, 0, 0), (0, 0, 2 * extent), n, n)
    noise = fx.RandomSampler(np.arange(len(v)), seed).get_float()
    v = v.copy()
    v[:, 1] = (0.4 * np.sin(1.1 * v[:, 0]) * np.cos(0.8 * v[:, 2]) + 0.25 * noise).astype(np.float32)
    return v.astype(np.float32), q


def case_quads(flags=0):
    v, q = bumpy_quads(28, 4.0)
    q = q.copy()
    q[17, 2] = 10 ** 6                                             # out-of-range index: the whole quad is dropped
    v2, q2 = bumpy_quads(6, 1.0, seed=9)
    v2 = (v2 * np.float32(0.8) + np.array([0.5, 1.6, -0.3], dtype=np.float32)).astype(np.float32)
    v2[10, 1] = np.nan                                             # the quads around this vertex vanish, both halves
    meshes = [(v, q), fx.triangle_sphere((-1.0, 1.2, 0.8), 0.7, 14), (v2, q2)]
    rays = np.concatenate([fx.incoherent_rays(7000, org=(0.1, 2.6, 0.2), seed=15),
                           fx.primary_rays(80, 80, org=(0.2, 7.0, 0.1), look=(0, -1, 0), up=(0, 0, 1)),
                           fx.primary_rays(40, 40, org=(0.0, -5.0, 0.3), look=(0, 1, 0), up=(0, 0, 1))])   # back faces
    rays["tnear"][::9] = 1e-3
    return dict(meshes=meshes, flags=flags, rays=rays)


CASES = {"quads_mixed": lambda: case_quads(0), "quads_mixed_robust": lambda: case_quads(rt.RTC_SCENE_FLAG_ROBUST)}


def load_golden(name):
    import os
    z = np.load(os.path.join(cases.ROOT, "tests", "golden", name + ".npz"))
    return dict(rays=z["rays_in"].view(rt.RAYHIT_DTYPE).reshape(-1), closest=z["closest_out"].view(rt.RAYHIT_DTYPE).reshape(-1),
                shadow_in=z["shadow_in"].view(rt.RAY_DTYPE).reshape(-1), shadow_out=z["shadow_out"].view(rt.RAY_DTYPE).reshape(-1),
                bounds=z["bounds_ref"])
