"""Synthetic scenes and ray streams of the BASELINE.json configurations (numpy, host side).

Restated from the reference's own benchmark fixtures so both libraries are fed identical input:
  triangle_plane / triangle_sphere   tutorials/common/scenegraph/geometry_creation.cpp:8-36,121-176
  RandomSampler (Murmur3 + LCG)      tutorials/common/math/random_sampler.h:15-118
  cosine_sample_hemisphere           tutorials/common/math/sampling.h:52-58
  incoherent benchmark rays          tutorials/verify/verify.cpp:4867-5016 (dir = 2*rand3-1 from a point)
Nothing here is on the product's compute path; it only makes inputs.
"""
import numpy as np

from .rtcore import RAYHIT_DTYPE, RAY_DTYPE, new_rays

U32 = np.uint32


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
def triangle_plane(p0, dx, dy, width, height):
    p0, dx, dy = (np.asarray(a, dtype=np.float32) for a in (p0, dx, dy))
    xs = (np.arange(width + 1, dtype=np.float32) / np.float32(width))
    ys = (np.arange(height + 1, dtype=np.float32) / np.float32(height))
    v = (p0[None, None, :] + xs[None, :, None] * dx[None, None, :] + ys[:, None, None] * dy[None, None, :])
    v = v.reshape(-1, 3).astype(np.float32)
    y, x = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    p00 = (y * (width + 1) + x).ravel()
    p01 = p00 + 1
    p10 = p00 + (width + 1)
    p11 = p10 + 1
    t = np.empty((2 * width * height, 3), dtype=np.uint32)
    t[0::2] = np.stack([p00, p01, p10], axis=1)
    t[1::2] = np.stack([p11, p10, p01], axis=1)
    return v, t


def quad_plane(p0, dx, dy, width, height):
    """Same grid as triangle_plane, one RTC_GEOMETRY_TYPE_QUAD per cell (v0..v3 counter-clockwise: p00, p01, p11, p10)."""
    v, _ = triangle_plane(p0, dx, dy, width, height)
    y, x = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    p00 = (y * (width + 1) + x).ravel()
    q = np.stack([p00, p00 + 1, p00 + (width + 1) + 1, p00 + (width + 1)], axis=1).astype(np.uint32)
    return v, q


def triangle_sphere(center, radius, num_phi):
    num_theta = 2 * num_phi
    phi = np.arange(num_phi + 1, dtype=np.float32) * np.float32(np.pi) * np.float32(1.0 / num_phi)
    theta = np.arange(num_theta, dtype=np.float32) * np.float32(2.0) * np.float32(np.pi) * np.float32(1.0 / num_theta)
    sp, cp = np.sin(phi)[:, None], np.cos(phi)[:, None]
    st, ct = np.sin(theta)[None, :], np.cos(theta)[None, :]
    v = np.empty((num_phi + 1, num_theta, 3), dtype=np.float32)
    v[..., 0] = center[0] + radius * sp * st
    v[..., 1] = center[1] + radius * cp * np.ones_like(st)
    v[..., 2] = center[2] + radius * sp * ct
    v = v.reshape(-1, 3)
    tris = []
    th = np.arange(1, num_theta + 1, dtype=np.int64)
    for p in range(1, num_phi + 1):
        if p == 1:
            p00 = np.full_like(th, num_theta - 1)
            p10 = p * num_theta + th - 1
            p11 = p * num_theta + th % num_theta
            tris.append(np.stack([p10, p00, p11], axis=1))
        elif p == num_phi:
            p00 = (p - 1) * num_theta + th - 1
            p01 = (p - 1) * num_theta + th % num_theta
            p10 = np.full_like(th, num_phi * num_theta)
            tris.append(np.stack([p10, p00, p01], axis=1))
        else:
            p00 = (p - 1) * num_theta + th - 1
            p01 = (p - 1) * num_theta + th % num_theta
            p10 = p * num_theta + th - 1
            p11 = p * num_theta + th % num_theta
            a = np.stack([p10, p00, p11], axis=1)
            b = np.stack([p01, p11, p00], axis=1)
            tris.append(np.stack([a, b], axis=1).reshape(-1, 3))
    return v, np.concatenate(tris).astype(np.uint32)


def _murmur_mix(h, k):
    k = (k * U32(0xcc9e2d51)).astype(U32)
    k = ((k << U32(15)) | (k >> U32(17))).astype(U32)
    k = (k * U32(0x1b873593)).astype(U32)
    h = (h ^ k).astype(U32)
    h = ((h << U32(13)) | (h >> U32(19))).astype(U32)
    return (h * U32(5) + U32(0xe6546b64)).astype(U32)


def _murmur_final(h):
    h = (h ^ (h >> U32(16))).astype(U32)
    h = (h * U32(0x85ebca6b)).astype(U32)
    h = (h ^ (h >> U32(13))).astype(U32)
    h = (h * U32(0xc2b2ae35)).astype(U32)
    return (h ^ (h >> U32(16))).astype(U32)


class RandomSampler:
    """Vectorised RandomSampler: one independent stream per element."""

    def __init__(self, pixel_id, sample_id=None):
        with np.errstate(over="ignore"):
            h = np.zeros(np.shape(pixel_id), dtype=U32)
            h = _murmur_mix(h, np.asarray(pixel_id).astype(U32))
            if sample_id is not None:
                h = _murmur_mix(h, np.broadcast_to(np.asarray(sample_id).astype(U32), h.shape).copy())
            self.s = _murmur_final(h)

    @classmethod
    def from_xy(cls, x, y, sample_id):
        return cls(np.asarray(x).astype(U32) | (np.asarray(y).astype(U32) << U32(16)), sample_id)

    def get_uint(self):
        with np.errstate(over="ignore"):
            self.s = (self.s * U32(1664525) + U32(1013904223)).astype(U32)
        return self.s

    def get_float(self):
        return (self.get_uint() >> U32(1)).astype(np.float32) * np.float32(4.656612873077392578125e-10)


def displaced_plane(n, extent=10.0, seed=1):
    """(n x n)-cell plane over [-extent,extent]^2 in xz, y = 0.5 sin(0.9x) cos(0.7z) + 0.05 noise."""
    v, t = triangle_plane((-extent, 0, -extent), (2 * extent, 0, 0), (0, 0, 2 * extent), n, n)
    noise = RandomSampler(np.arange(len(v)), seed).get_float()
    v[:, 1] = (0.5 * np.sin(0.9 * v[:, 0]) * np.cos(0.7 * v[:, 2]) + 0.05 * noise).astype(np.float32)
    return v, t


# ------------------------------------------------------------------------------------------------
# the named scenes
# ------------------------------------------------------------------------------------------------
def scene_c1():
    """config 0: tessellated unit sphere, numPhi=91 -> 32 760 triangles."""
    return [triangle_sphere((0.0, 0.0, 0.0), 1.0, 91)]


def scene_c2(scale=1.0):
    """config 1: 700x700 displaced plane (980 000 tris) + sphere numPhi=72 (20 448 tris) ~ 1.0 M tris.
    scale < 1 shrinks the tessellation for tests."""
    n = max(2, int(round(700 * scale)))
    s = max(3, int(round(72 * scale)))
    return [displaced_plane(n), triangle_sphere((0.0, 2.0, 0.0), 1.5, s)]


def scene_c3(scale=1.0):
    """config 2-4: 2200x2200 displaced plane (9.68 M) + 4 spheres numPhi=142 (4 x 80 088) ~ 10.0 M tris."""
    n = max(2, int(round(2200 * scale)))
    s = max(3, int(round(142 * scale)))
    cs = [(-4.0, 2.0, -4.0), (4.0, 2.0, -4.0), (-4.0, 2.0, 4.0), (4.0, 2.0, 4.0)]
    return [displaced_plane(n)] + [triangle_sphere(c, 1.5, s) for c in cs]


def random_soup(n, seed=7):
    """config 4 stress case: n small triangles uniform in the unit cube, edge ~ n^(-1/3)."""
    rs = RandomSampler(np.arange(n), seed)
    c = np.stack([rs.get_float() for _ in range(3)], axis=1)
    e = np.float32(n ** (-1.0 / 3.0))
    o = [np.stack([rs.get_float() for _ in range(3)], axis=1) * e for _ in range(3)]
    v = np.stack([c + o[0] - e / 2, c + o[1] - e / 2, c + o[2] - e / 2], axis=1).reshape(-1, 3).astype(np.float32)
    t = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    return [(v, t)]


def num_tris(meshes):
    """Triangles the builder sees: a quad (4-index record) counts as two."""
    return int(sum(len(t) * (2 if (np.ndim(t) == 2 and np.shape(t)[1] == 4) else 1) for _, t in meshes))


# ------------------------------------------------------------------------------------------------
# ray streams
# ------------------------------------------------------------------------------------------------
def _set(rays, org, dirs, tnear, tfar):
    rays["org_x"], rays["org_y"], rays["org_z"] = org[..., 0], org[..., 1], org[..., 2]
    rays["dir_x"], rays["dir_y"], rays["dir_z"] = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    rays["tnear"] = tnear
    rays["tfar"] = tfar
    return rays


def primary_rays(width, height, org, look, up=(0, 0, 1), fov_scale=1.0, hit=True, rows=None):
    """Pinhole camera, row-major pixel order, unnormalised-then-normalised directions (coherent)."""
    org = np.asarray(org, dtype=np.float32)
    w = np.asarray(look, dtype=np.float32)
    w = w / np.linalg.norm(w)
    u = np.cross(np.asarray(up, dtype=np.float32), w)
    u = (u / np.linalg.norm(u)).astype(np.float32)
    v = np.cross(w, u).astype(np.float32)
    r0, r1 = rows if rows is not None else (0, height)
    ys, xs = np.meshgrid(np.arange(r0, r1, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    sx = ((xs + 0.5) / width - 0.5) * np.float32(fov_scale)
    sy = ((ys + 0.5) / height - 0.5) * np.float32(fov_scale)
    d = sx[..., None] * u + sy[..., None] * v + w
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32).reshape(-1, 3)
    rays = new_rays(len(d), hit)
    rays["id"] = np.arange(r0 * width, r1 * width, dtype=np.uint32)
    return _set(rays, np.broadcast_to(org, d.shape), d, 0.0, np.inf)


def incoherent_rays(n, org=(0.0, 0.0, 0.0), seed=0, hit=True, first=0):
    """The reference's IncoherentRaysBenchmark: dir = 2*rand3 - 1 from one point (not normalised)."""
    rs = RandomSampler(np.arange(first, first + n), seed)
    d = np.stack([(rs.get_uint() >> U32(1)).astype(np.float32) * np.float32(4.656612873077392578125e-10)
                  for _ in range(3)], axis=1)
    d = (2.0 * d - 1.0).astype(np.float32)
    rays = new_rays(n, hit)
    rays["id"] = np.arange(first, first + n, dtype=np.uint32)
    return _set(rays, np.broadcast_to(np.asarray(org, dtype=np.float32), d.shape), d, 0.0, np.inf)


def cosine_sample_hemisphere(sx, sy):
    phi = np.float32(2.0 * np.pi) * sx
    cos_t = np.sqrt(sy)
    sin_t = np.sqrt(np.float32(1.0) - sy)
    return np.stack([np.cos(phi) * sin_t, np.sin(phi) * sin_t, cos_t], axis=-1).astype(np.float32)


def _frame(n):
    """Orthonormal frame (dx,dy,n) per normal."""
    a = np.where(np.abs(n[:, 0:1]) > 0.9, np.array([[0, 1, 0]], dtype=np.float32), np.array([[1, 0, 0]], dtype=np.float32))
    dx = np.cross(a, n)
    dx /= np.linalg.norm(dx, axis=1, keepdims=True)
    dy = np.cross(n, dx)
    return dx.astype(np.float32), dy.astype(np.float32)


def hit_points(rays):
    """(P, Ng face-forwarded & normalised, hit mask) of a traced RTCRayHit stream."""
    ok = rays["geomID"] != 0xFFFFFFFF
    o = np.stack([rays["org_x"], rays["org_y"], rays["org_z"]], axis=1)
    d = np.stack([rays["dir_x"], rays["dir_y"], rays["dir_z"]], axis=1)
    n = np.stack([rays["Ng_x"], rays["Ng_y"], rays["Ng_z"]], axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.sum(n * d, axis=1) > 0
    n[flip] = -n[flip]
    p = o + rays["tfar"][:, None] * d
    return p.astype(np.float32), n.astype(np.float32), ok


def diffuse_rays(traced, sample_id=0, hit=True):
    """Secondary stream: cosine-weighted bounce from every hit of `traced` (pixel order kept =>
    incoherent directions), origin offset 1e-3 along the normal, tnear 1e-3, tfar inf."""
    p, n, ok = hit_points(traced)
    p, n = p[ok], n[ok]
    ids = traced["id"][ok]
    rs = RandomSampler(ids, sample_id)
    sx, sy = rs.get_float(), rs.get_float()
    l = cosine_sample_hemisphere(sx, sy)
    dx, dy = _frame(n)
    d = (l[:, 0:1] * dx + l[:, 1:2] * dy + l[:, 2:3] * n).astype(np.float32)
    rays = new_rays(len(p), hit)
    rays["id"] = ids
    return _set(rays, (p + np.float32(1e-3) * n).astype(np.float32), d, 1e-3, np.inf)


def shadow_rays(traced, light=(5.0, 10.0, 5.0)):
    """Occlusion stream from every hit towards a point light, tfar = distance*(1-1e-4) (dir normalised)."""
    p, n, ok = hit_points(traced)
    p, n = p[ok], n[ok]
    o = (p + np.float32(1e-3) * n).astype(np.float32)
    d = np.asarray(light, dtype=np.float32)[None, :] - o
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    rays = new_rays(len(o), hit=False)
    rays["id"] = traced["id"][ok]
    return _set(rays, o, d, 1e-3, dist * np.float32(1.0 - 1e-4))


def to_ray(rayhit):
    """RTCRay view-copy (first 48 bytes) of an RTCRayHit stream."""
    out = np.zeros(len(rayhit), dtype=RAY_DTYPE)
    for k in RAY_DTYPE.names:
        out[k] = rayhit[k]
    return out


# Cameras used by the configurations (all primary rays hit the ground plane).
C1_CAMERA = dict(org=(0.0, 0.0, -3.0), look=(0.0, 0.0, 1.0), up=(0, 1, 0), fov_scale=1.0)
C2_CAMERA = dict(org=(0.0, 15.0, 0.0), look=(0.0, -1.0, 0.0), up=(0, 0, 1), fov_scale=1.2)
