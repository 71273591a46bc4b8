"""Deterministic small parity cases shared by the golden generator (tests/golden/make_golden.py),
the oracle tests and the GPU parity tests.  Each case = meshes + scene flags + one RTCRayHit stream.
The catalogue follows the reference's own triangle-path tests (tutorials/verify/verify.cpp):
TriangleHitTest :2339-2426, WatertightTest :2898-2979, SmallTriangleHitTest :2981-3048,
InactiveRaysTest :2838-2896, NaNTest :3101 / InfTest :3174, GarbageGeometryTest :1792,
OverlappingGeometryTest :1216, BufferStrideTest :922.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore

ROBUST = rt.RTC_SCENE_FLAG_ROBUST


def _rays(org, dirs, tnear=0.0, tfar=np.inf):
    org = np.broadcast_to(np.asarray(org, dtype=np.float32), np.shape(dirs)).astype(np.float32)
    r = rt.new_rays(len(dirs))
    r["id"] = np.arange(len(dirs), dtype=np.uint32)
    return fx._set(r, org, np.asarray(dirs, dtype=np.float32), tnear, tfar)


def case_triangle_hit():
    """TriangleHitTest: triangle (0,0,0),(1,0,0),(0,1,0); 256 rays from (0,0,-1) to sampled points.
    Returns expected (u,v) as well: |u-u0|,|v-v0|,|t-1| <= 16 ulp, Ng = (0,0,1)."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    t = np.array([[0, 1, 2]], dtype=np.uint32)
    rs = fx.RandomSampler(np.arange(256), 11)
    u, w = rs.get_float(), rs.get_float()
    su = np.sqrt(u).astype(np.float32)
    w = (w * su).astype(np.float32)
    u = (np.float32(1.0) - su).astype(np.float32)
    bad = (u < 0.001) | (w < 0.001) | ((u + w) > 0.999)
    u[bad], w[bad] = 0.333, 0.333
    a, b, c = v[1], v[2], v[0]                       # uniformSampleTriangle(vertices[1], vertices[2], vertices[0])
    to = c[None] + u[:, None] * (a - c)[None] + w[:, None] * (b - c)[None]
    org = np.array([0, 0, -1], dtype=np.float32)
    return dict(meshes=[(v, t)], flags=0, rays=_rays(org, (to - org).astype(np.float32)), expect_uv=(u, w))


def case_sphere_small():
    meshes = [fx.triangle_sphere((0.0, 0.0, 0.0), 1.0, 24)]
    inside = fx.incoherent_rays(4096, org=(0.1, 0.2, -0.1), seed=21)
    outside = fx.primary_rays(64, 64, org=(0.0, 0.0, -3.0), look=(0, 0, 1), up=(0, 1, 0))
    return dict(meshes=meshes, flags=0, rays=np.concatenate([inside, outside]))


def case_two_geoms():
    meshes = [fx.displaced_plane(30, extent=4.0), fx.triangle_sphere((0.0, 1.5, 0.0), 1.0, 12)]
    rs = fx.RandomSampler(np.arange(6000), 31)
    o = np.stack([rs.get_float() * 6 - 3, rs.get_float() * 2 + 0.8, rs.get_float() * 6 - 3], 1).astype(np.float32)
    d = np.stack([rs.get_float() * 2 - 1, rs.get_float() * 2 - 1.3, rs.get_float() * 2 - 1], 1).astype(np.float32)
    return dict(meshes=meshes, flags=0, rays=_rays(o, d, 1e-3, np.inf))


def case_robust_far_sphere():
    """WatertightTest geometry: sphere far from the origin, rays from its centre must all hit."""
    c = (148376.0, 1234.0, -223423.0)
    meshes = [fx.triangle_sphere(c, 2.0, 20)]
    return dict(meshes=meshes, flags=ROBUST, rays=fx.incoherent_rays(4096, org=c, seed=41))


def case_small_triangles():
    """SmallTriangleHitTest: rays aimed at triangle centroids of a fine plane must report that primID."""
    v, t = fx.triangle_plane((-1, -1, 0), (2, 0, 0), (0, 2, 0), 40, 40)
    ang = np.float32(0.3)
    rot = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype=np.float32)
    v = (v @ rot.T).astype(np.float32)
    rs = fx.RandomSampler(np.arange(2000), 51)
    pick = (rs.get_uint() % np.uint32(len(t))).astype(np.int64)
    cen = ((v[t[pick, 0]] + v[t[pick, 1]] + v[t[pick, 2]]) / np.float32(3.0)).astype(np.float32)
    org = np.array([0.2, -0.1, -5.0], dtype=np.float32)
    return dict(meshes=[(v, t)], flags=0, rays=_rays(org, (cen - org).astype(np.float32)), expect_prim=pick.astype(np.uint32))


def case_edge_rays():
    """Inactive / degenerate / special-value rays over the small sphere (stream semantics)."""
    meshes = [fx.triangle_sphere((0.0, 0.0, 0.0), 1.0, 16)]
    base = fx.incoherent_rays(512, org=(0.05, 0.0, 0.1), seed=61)
    r = base.copy()
    n = len(r)
    k = np.arange(n)
    r["tnear"][k % 16 == 1] = np.inf; r["tfar"][k % 16 == 1] = 0.0          # InactiveRaysTest
    r["tnear"][k % 16 == 2] = 2.0; r["tfar"][k % 16 == 2] = 1.0             # tnear > tfar
    r["tfar"][k % 16 == 3] = 0.5                                            # segment ends before the surface
    r["tnear"][k % 16 == 4] = 5.0                                           # segment starts behind the surface
    r["dir_x"][k % 16 == 5] = 0.0                                           # axis-parallel components
    r["dir_y"][k % 16 == 6] = 0.0; r["dir_z"][k % 16 == 6] = 0.0
    r["org_x"][k % 16 == 7] = np.nan                                        # NaNTest
    r["dir_y"][k % 16 == 8] = np.nan
    r["tfar"][k % 16 == 9] = np.nan
    r["org_z"][k % 16 == 10] = np.inf                                       # InfTest
    r["dir_x"][k % 16 == 11] = -np.inf
    r["tnear"][k % 16 == 12] = -1.0                                         # negative tnear (clamped for boxes)
    r["tfar"][k % 16 == 13] = -1.0                                          # tfar < 0
    for f in ("dir_x", "dir_y", "dir_z"):
        r[f][k % 16 == 14] = 0.0                                            # zero direction
    r["dir_x"][k % 16 == 15] *= np.float32(1e-20)                           # |dir| below min_rcp_input
    # occlusion streams are traced per category: the reference picks the near/far planes of a whole
    # 32-ray octant chunk from its FIRST ray (bvh_intersector_stream.cpp:324), so one NaN ray in a
    # stream corrupts the answers of its valid neighbours; categories must not share a stream.
    return dict(meshes=meshes, flags=0, rays=r, groups=(k % 16).astype(np.int32))


def case_garbage_prims():
    """GarbageGeometryTest flavour: out-of-range indices and non-finite / huge vertices are dropped."""
    v, t = fx.triangle_sphere((0.0, 0.0, 0.0), 1.0, 12)
    v, t = v.copy(), t.copy()
    t[5] = (0, 1, 10 ** 6)                # index out of range
    t[17] = (0xFFFFFFFF, 2, 3)
    v[40] = (np.nan, 0, 0)                # every triangle touching vertex 40 / 77 / 120 disappears
    v[77] = (0, np.inf, 0)
    v[120] = (3e18, 0, 0)                 # beyond FLT_LARGE = 1.844e18
    return dict(meshes=[(v, t)], flags=0, rays=fx.incoherent_rays(4096, org=(0.0, 0.1, 0.0), seed=71))


def case_overlapping():
    """OverlappingGeometryTest flavour: the same mesh attached twice (coincident triangles, geomID 0 and 1)."""
    m = fx.triangle_sphere((0.0, 0.0, 0.0), 1.0, 10)
    return dict(meshes=[m, (m[0].copy(), m[1].copy())], flags=0, rays=fx.incoherent_rays(2048, org=(0, 0, 0), seed=81))


CASES = {
    "triangle_hit": case_triangle_hit,
    "sphere_small": case_sphere_small,
    "two_geoms": case_two_geoms,
    "robust_far_sphere": case_robust_far_sphere,
    "small_triangles": case_small_triangles,
    "edge_rays": case_edge_rays,
    "garbage_prims": case_garbage_prims,
    "overlapping": case_overlapping,
}

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Inputs and reference outputs exactly as stored by make_golden.py (no regeneration => no libm drift)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    nm = int(z["num_meshes"])
    meshes = [(z[f"v{i}"], z[f"t{i}"]) for i in range(nm)]
    out = dict(meshes=meshes, flags=int(z["flags"]), rays=z["rays_in"].view(rt.RAYHIT_DTYPE).reshape(-1),
               closest=z["closest_out"].view(rt.RAYHIT_DTYPE).reshape(-1),
               shadow_in=z["shadow_in"].view(rt.RAY_DTYPE).reshape(-1),
               shadow_out=z["shadow_out"].view(rt.RAY_DTYPE).reshape(-1),
               occl_self_out=z["occl_self_out"].view(rt.RAY_DTYPE).reshape(-1),
               sah_ref=float(z["sah_ref"]), bounds_ref=z["bounds_ref"])
    for k in ("expect_u", "expect_v", "expect_prim", "groups"):
        if k in z.files:
            out[k] = z[k]
    return out


def occluded_by_group(trace, rays, groups):
    """Apply trace(stream) to the whole stream, or once per category when the case defines groups."""
    if groups is None:
        trace(rays)
        return rays
    for gid in np.unique(groups):
        part = rays[groups == gid].copy()
        trace(part)
        rays[groups == gid] = part
    return rays
