"""ctypes wrapper of oracle/librq_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "librq_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libembree3_ref.so")


class _Mesh(C.Structure):
    _fields_ = [("indices", C.c_void_p), ("vertices", C.c_void_p), ("indexStride", C.c_uint), ("vertexStride", C.c_uint),
                ("numTris", C.c_uint), ("numVerts", C.c_uint), ("geomID", C.c_uint), ("quads", C.c_uint)]


class _Instance(C.Structure):
    _fields_ = [("scene", C.c_void_p), ("l2w", C.c_float * 12), ("geomID", C.c_uint)]


def build_lib():
    subprocess.check_call(["make", "-s", "-C", HERE, "librq_oracle.so"])
    return LIB


class Oracle:
    def __init__(self):
        if not os.path.exists(LIB):
            build_lib()
        self.lib = L = C.CDLL(LIB)
        L.rqo_build.restype = C.c_void_p
        L.rqo_build.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_int]
        L.rqo_free.argtypes = [C.c_void_p]
        L.rqo_sah.restype = C.c_double
        L.rqo_sah.argtypes = [C.c_void_p]
        L.rqo_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rqo_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.rqo_intersect1M.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_uint]
        L.rqo_occluded1M.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t]
        L.rqo_build_top.restype = C.c_void_p
        L.rqo_build_top.argtypes = [C.c_void_p, C.POINTER(_Instance), C.c_int]
        L.rqo_free_top.argtypes = [C.c_void_p]
        L.rqo_top_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.rqo_top_intersect1M.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_uint]
        L.rqo_top_occluded1M.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t]
        fp = C.POINTER(C.c_float)
        for n in ("rqo_moeller", "rqo_pluecker"):
            f = getattr(L, n)
            f.restype = C.c_int
            f.argtypes = [fp, fp, C.c_float, C.c_float, fp, fp, fp, fp]

    def build(self, meshes, robust=False, geom_ids=None):
        """meshes: list of (vertices (n,3) f32, triangles (m,3) u32) -- or quads (m,4) u32: RTC_GEOMETRY_TYPE_QUAD."""
        arr = (_Mesh * max(len(meshes), 1))()
        keep = []
        for i, (v, t) in enumerate(meshes):
            v = np.ascontiguousarray(v, dtype=np.float32)
            t = np.ascontiguousarray(t, dtype=np.uint32)
            keep += [v, t]
            quads = 1 if (t.ndim == 2 and t.shape[1] == 4) else 0
            arr[i] = _Mesh(t.ctypes.data, v.ctypes.data, 16 if quads else 12, 12, len(t), len(v), geom_ids[i] if geom_ids else i, quads)
        h = self.lib.rqo_build(arr, len(meshes), 1 if robust else 0)
        return h

    def free(self, h):
        self.lib.rqo_free(h)

    # ---- single-level instancing: a top-level scene = optional triangle scene + instances of other scenes ----
    def build_top(self, base, instances):
        """base: handle from build() or None; instances: list of (scene handle, l2w (12,) column major vx,vy,vz,p, geomID)."""
        arr = (_Instance * max(len(instances), 1))()
        for i, (h, m, gid) in enumerate(instances):
            arr[i].scene = h
            arr[i].l2w = (C.c_float * 12)(*[float(x) for x in np.asarray(m, dtype=np.float32).ravel()])
            arr[i].geomID = gid
        return self.lib.rqo_build_top(base, arr, len(instances))

    def free_top(self, h):
        self.lib.rqo_free_top(h)

    def top_bounds(self, h):
        o = (C.c_float * 6)()
        self.lib.rqo_top_bounds(h, o)
        return np.array(list(o), dtype=np.float32)

    def top_intersect(self, h, rays, inst_id=0xFFFFFFFF):
        self.lib.rqo_top_intersect1M(h, rays.ctypes.data, len(rays), rays.strides[0], inst_id)

    def top_occluded(self, h, rays):
        self.lib.rqo_top_occluded1M(h, rays.ctypes.data, len(rays), rays.strides[0])

    def sah(self, h):
        return self.lib.rqo_sah(h)

    def stats(self, h):
        o = (C.c_uint64 * 4)()
        self.lib.rqo_stats(h, o)
        return dict(tris=o[0], nodes=o[1], leaves=o[2], blocks=o[3])

    def trace_counters(self, reset=True):
        """Work of the restated reference traversal since the last reset: rays, inner nodes, leaves, Triangle4 blocks, triangles."""
        o = (C.c_uint64 * 5)()
        self.lib.rqo_trace_counters(o, 1 if reset else 0)
        return dict(rays=o[0], nodes=o[1], leaves=o[2], blocks=o[3], tris=o[4])

    def bounds(self, h):
        o = (C.c_float * 6)()
        self.lib.rqo_bounds(h, o)
        return np.array(list(o), dtype=np.float32)

    def intersect(self, h, rays, inst_id=0xFFFFFFFF):
        self.lib.rqo_intersect1M(h, rays.ctypes.data, len(rays), rays.strides[0], inst_id)

    def occluded(self, h, rays):
        self.lib.rqo_occluded1M(h, rays.ctypes.data, len(rays), rays.strides[0])

    def tri_test(self, org, dir, tnear, tfar, v0, v1, v2, robust=False):
        f = self.lib.rqo_pluecker if robust else self.lib.rqo_moeller
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (org, dir, v0, v1, v2)]
        out = np.zeros(6, dtype=np.float32)
        p = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        ok = f(p(a[0]), p(a[1]), tnear, tfar, p(a[2]), p(a[3]), p(a[4]), p(out))
        return bool(ok), out
