"""Single-level instancing parity cases, shared by the golden generator
(tests/golden/make_golden_instances.py), the oracle tests and the GPU tests.  Reference behaviour:
kernels/geometry/instance_intersector.cpp:48-105 -- hits carry the instanced scene's geomID/primID,
Ng in instance space, instID[0] = geomID of the instance geometry in the top-level scene."""
import numpy as np

import cases

fx, rt = cases.fx, cases.rt


def _xfm(angle_y, angle_x, scale, trans):
    """local2world, column major (vx, vy, vz, p), float32: R_y(angle_y) R_x(angle_x) diag(scale) then translation."""
    cy, sy, cx, sx = np.cos(angle_y), np.sin(angle_y), np.cos(angle_x), np.sin(angle_x)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    m = ry @ rx @ np.diag(scale)
    return np.concatenate([m[:, 0], m[:, 1], m[:, 2], np.asarray(trans, dtype=np.float64)]).astype(np.float32)


def case_forest(flags=0):
    objects = [[fx.triangle_sphere((0.0, 0.0, 0.0), 0.5, 14)],
               [fx.displaced_plane(10, extent=0.8), fx.triangle_sphere((0.2, 0.5, 0.1), 0.3, 8)]]   # object 1 holds two geometries
    base = [fx.displaced_plane(20, extent=6.0)]
    rs = fx.RandomSampler(np.arange(14), 77)
    inst = []
    for i in range(14):
        a, b = float(rs.get_float()[i]) * 6.28, float(rs.get_float()[i]) * 1.2 - 0.6
        sc = [0.6 + 1.2 * float(rs.get_float()[i]), 0.6 + 1.0 * float(rs.get_float()[i]), 0.6 + 1.4 * float(rs.get_float()[i])]
        if i == 5:
            sc[0] = -sc[0]                                         # mirrored instance (negative determinant)
        tr = [float(rs.get_float()[i]) * 9 - 4.5, 0.6 + 1.6 * float(rs.get_float()[i]), float(rs.get_float()[i]) * 9 - 4.5]
        inst.append((i % 2, _xfm(a, b, sc, tr)))
    inst.append((0, _xfm(0.0, 0.0, [1, 1, 1], [0.0, 1.0, 0.0])))   # identity-rotation instance at a known place
    inst.append((0, _xfm(0.3, 0.1, [1.2, 0.9, 1.1], [0.1, 1.1, 0.1])))   # overlaps the previous one
    rays = np.concatenate([fx.incoherent_rays(6000, org=(0.05, 3.0, 0.1), seed=5),
                           fx.incoherent_rays(3000, org=(0.0, 1.0, 0.0), seed=6),          # starts inside an instance
                           fx.primary_rays(72, 72, org=(0.3, 9.0, 0.2), look=(0, -1, 0), up=(0, 0, 1))])
    rays["tnear"][::7] = 1e-3
    rays["tfar"][::11] = 4.0
    rays["tfar"][5::97] = -1.0                                     # inactive
    return dict(objects=objects, base=base, instances=inst, flags=flags, rays=rays)


def case_instances_only():
    """No triangle geometry in the top-level scene; one instanced scene, a grid of translated copies."""
    objects = [[fx.triangle_sphere((0.0, 0.0, 0.0), 0.4, 10)]]
    inst = [(0, _xfm(0.1 * (x + z), 0.0, [1.0, 1.0 + 0.1 * x, 1.0], [1.5 * x, 0.0, 1.5 * z])) for x in range(-2, 3) for z in range(-2, 3)]
    rays = np.concatenate([fx.primary_rays(64, 64, org=(0.2, 7.0, 0.1), look=(0, -1, 0), up=(0, 0, 1)),
                           fx.incoherent_rays(4000, org=(0.7, 0.1, 0.8), seed=8)])
    return dict(objects=objects, base=[], instances=inst, flags=0, rays=rays)


CASES = {"inst_forest": lambda: case_forest(0), "inst_forest_robust": lambda: case_forest(rt.RTC_SCENE_FLAG_ROBUST),
         "inst_only": case_instances_only}


def build_oracle(oracle, c):
    """-> (top handle, list of scene handles to free)."""
    robust = bool(c["flags"] & rt.RTC_SCENE_FLAG_ROBUST)
    objs = [oracle.build(m, robust=robust) for m in c["objects"]]
    base = oracle.build(c["base"], robust=robust) if c["base"] else None
    nbase = len(c["base"])
    top = oracle.build_top(base, [(objs[oi], m, nbase + i) for i, (oi, m) in enumerate(c["instances"])])
    return top, objs + ([base] if base else [])


def load_golden(name):
    import os
    z = np.load(os.path.join(cases.ROOT, "tests", "golden", name + ".npz"))
    return dict(rays=z["rays_in"].view(rt.RAYHIT_DTYPE).reshape(-1), closest=z["closest_out"].view(rt.RAYHIT_DTYPE).reshape(-1),
                shadow_in=z["shadow_in"].view(rt.RAY_DTYPE).reshape(-1), shadow_out=z["shadow_out"].view(rt.RAY_DTYPE).reshape(-1),
                bounds=z["bounds_ref"])
