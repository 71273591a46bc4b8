"""ctypes binding of the rtcore C ABI (include/embree3/rtcore.h, include/rq_b200.h).

The same binding drives any library that exports the `rtc*` symbols: the product
(`embree-aarch64_b200/lib/libembree3.so`) and, in tests and the CPU baseline only, a build of the
reference library (path supplied by the caller; this package never looks for it).  That is the point of the drop-in boundary: one harness,
two libraries, identical calls (the shape of the reference's own `IntersectWithMode`,
tutorials/verify/rtcore_helpers.h:751-879).

Ray streams are numpy structured arrays with the exact AoS layouts of rtcore_ray.h:11-49
(`RAY_DTYPE` 48 B, `RAYHIT_DTYPE` 80 B) or raw device pointers (ints) for GPU-resident streams.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "lib", "libembree3.so")

RTC_INVALID_GEOMETRY_ID = 0xFFFFFFFF
RTC_FORMAT_UINT3 = 0x5003
RTC_FORMAT_UINT4 = 0x5004
RTC_FORMAT_FLOAT3 = 0x9003
RTC_BUFFER_TYPE_INDEX = 0
RTC_BUFFER_TYPE_VERTEX = 1
RTC_GEOMETRY_TYPE_TRIANGLE = 0
RTC_GEOMETRY_TYPE_QUAD = 1
RTC_GEOMETRY_TYPE_INSTANCE = 121
RTC_FORMAT_FLOAT3X4_ROW_MAJOR = 0x9134
RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR = 0x9234
RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR = 0x9244
RTC_SCENE_FLAG_NONE = 0
RTC_SCENE_FLAG_DYNAMIC = 1
RTC_SCENE_FLAG_COMPACT = 2
RTC_SCENE_FLAG_ROBUST = 4
(RTC_BUILD_QUALITY_LOW, RTC_BUILD_QUALITY_MEDIUM, RTC_BUILD_QUALITY_HIGH, RTC_BUILD_QUALITY_REFIT) = range(4)
RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT = 0
RTC_INTERSECT_CONTEXT_FLAG_COHERENT = 1
(RTC_ERROR_NONE, RTC_ERROR_UNKNOWN, RTC_ERROR_INVALID_ARGUMENT, RTC_ERROR_INVALID_OPERATION,
 RTC_ERROR_OUT_OF_MEMORY, RTC_ERROR_UNSUPPORTED_CPU, RTC_ERROR_CANCELLED) = range(7)

RAY_FIELDS = [("org_x", "<f4"), ("org_y", "<f4"), ("org_z", "<f4"), ("tnear", "<f4"),
              ("dir_x", "<f4"), ("dir_y", "<f4"), ("dir_z", "<f4"), ("time", "<f4"),
              ("tfar", "<f4"), ("mask", "<u4"), ("id", "<u4"), ("flags", "<u4")]
HIT_FIELDS = [("Ng_x", "<f4"), ("Ng_y", "<f4"), ("Ng_z", "<f4"), ("u", "<f4"), ("v", "<f4"),
              ("primID", "<u4"), ("geomID", "<u4"), ("instID", "<u4")]
RAY_DTYPE = np.dtype(RAY_FIELDS)
RAYHIT_DTYPE = np.dtype(RAY_FIELDS + HIT_FIELDS)
assert RAY_DTYPE.itemsize == 48 and RAYHIT_DTYPE.itemsize == 80


class IntersectContext(C.Structure):
    _fields_ = [("flags", C.c_int), ("filter", C.c_void_p), ("instID", C.c_uint * 1)]


class Bounds(C.Structure):
    _fields_ = [("lower_x", C.c_float), ("lower_y", C.c_float), ("lower_z", C.c_float), ("align0", C.c_float),
                ("upper_x", C.c_float), ("upper_y", C.c_float), ("upper_z", C.c_float), ("align1", C.c_float)]


class BuildStats(C.Structure):
    _fields_ = [("numPrimsIn", C.c_uint), ("numPrimsValid", C.c_uint), ("numNodes", C.c_uint),
                ("numTris", C.c_uint), ("depth", C.c_uint), ("numLeaves", C.c_uint),
                ("sah", C.c_double), ("sahExact", C.c_double),
                ("msTotal", C.c_float), ("msPrims", C.c_float), ("msSort", C.c_float),
                ("msHierarchy", C.c_float), ("msRefit", C.c_float), ("msEmit", C.c_float),
                ("bytes", C.c_ulonglong), ("builderIterations", C.c_uint), ("refitCount", C.c_uint),
                ("sahInner", C.c_double), ("sahLeafTris", C.c_double), ("numTreelets", C.c_uint), ("msBroadcast", C.c_float), ("numSplitRefs", C.c_uint), ("pad", C.c_uint)]


class TraceCounters(C.Structure):
    _fields_ = [("rays", C.c_ulonglong), ("nodes", C.c_ulonglong), ("tris", C.c_ulonglong),
                ("hits", C.c_ulonglong), ("stackMax", C.c_ulonglong), ("emptyNodes", C.c_ulonglong),
                ("hitNodes", C.c_ulonglong), ("lateNodes", C.c_ulonglong)]


def new_rays(n, hit=True):
    """Zeroed stream with the conventions every caller must set up: geomID = instID = -1."""
    a = np.zeros(n, dtype=RAYHIT_DTYPE if hit else RAY_DTYPE)
    a["mask"] = 0xFFFFFFFF
    if hit:
        a["geomID"] = RTC_INVALID_GEOMETRY_ID
        a["primID"] = RTC_INVALID_GEOMETRY_ID
        a["instID"] = RTC_INVALID_GEOMETRY_ID
    return a


def _sig(lib, name, res, args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = args
    return f


class RTCore:
    """One loaded rtcore library."""

    def __init__(self, path=PRODUCT_LIB):
        if not os.path.exists(path):
            raise RuntimeError(f"rtcore library not built: {path} (run __graft_entry__.build())")
        self.path = path
        self.lib = L = C.CDLL(path)
        vp, u, sz = C.c_void_p, C.c_uint, C.c_size_t
        _sig(L, "rtcNewDevice", vp, [C.c_char_p])
        _sig(L, "rtcReleaseDevice", None, [vp])
        _sig(L, "rtcRetainDevice", None, [vp])
        _sig(L, "rtcGetDeviceError", C.c_int, [vp])
        _sig(L, "rtcGetDeviceProperty", C.c_ssize_t, [vp, C.c_int])
        _sig(L, "rtcNewScene", vp, [vp])
        _sig(L, "rtcReleaseScene", None, [vp])
        _sig(L, "rtcSetSceneFlags", None, [vp, C.c_int])
        _sig(L, "rtcGetSceneFlags", C.c_int, [vp])
        _sig(L, "rtcSetSceneBuildQuality", None, [vp, C.c_int])
        _sig(L, "rtcCommitScene", None, [vp])
        _sig(L, "rtcJoinCommitScene", None, [vp])
        _sig(L, "rtcGetSceneBounds", None, [vp, C.POINTER(Bounds)])
        _sig(L, "rtcNewGeometry", vp, [vp, C.c_int])
        _sig(L, "rtcReleaseGeometry", None, [vp])
        _sig(L, "rtcCommitGeometry", None, [vp])
        _sig(L, "rtcEnableGeometry", None, [vp])
        _sig(L, "rtcDisableGeometry", None, [vp])
        _sig(L, "rtcAttachGeometry", u, [vp, vp])
        _sig(L, "rtcAttachGeometryByID", None, [vp, vp, u])
        _sig(L, "rtcDetachGeometry", None, [vp, u])
        _sig(L, "rtcGetGeometry", vp, [vp, u])
        _sig(L, "rtcSetSharedGeometryBuffer", None, [vp, C.c_int, u, C.c_int, vp, sz, sz, sz])
        _sig(L, "rtcSetNewGeometryBuffer", vp, [vp, C.c_int, u, C.c_int, sz, sz])
        _sig(L, "rtcGetGeometryBufferData", vp, [vp, C.c_int, u])
        _sig(L, "rtcUpdateGeometryBuffer", None, [vp, C.c_int, u])
        _sig(L, "rtcSetGeometryTimeStepCount", None, [vp, u])
        _sig(L, "rtcSetGeometryBuildQuality", None, [vp, C.c_int])
        _sig(L, "rtcSetGeometryInstancedScene", None, [vp, vp])
        _sig(L, "rtcSetGeometryTransform", None, [vp, u, C.c_int, vp])
        _sig(L, "rtcGetGeometryTransform", None, [vp, C.c_float, C.c_int, vp])
        _sig(L, "rtcNewBuffer", vp, [vp, sz])
        _sig(L, "rtcNewSharedBuffer", vp, [vp, vp, sz])
        _sig(L, "rtcGetBufferData", vp, [vp])
        _sig(L, "rtcReleaseBuffer", None, [vp])
        _sig(L, "rtcSetGeometryBuffer", None, [vp, C.c_int, u, C.c_int, vp, sz, sz, sz])
        ctxp = C.POINTER(IntersectContext)
        _sig(L, "rtcIntersect1", None, [vp, ctxp, vp])
        _sig(L, "rtcOccluded1", None, [vp, ctxp, vp])
        _sig(L, "rtcIntersect1M", None, [vp, ctxp, vp, u, sz])
        _sig(L, "rtcOccluded1M", None, [vp, ctxp, vp, u, sz])
        _sig(L, "rtcIntersect1Mp", None, [vp, ctxp, vp, u])
        _sig(L, "rtcOccluded1Mp", None, [vp, ctxp, vp, u])
        _sig(L, "rtcIntersectNM", None, [vp, ctxp, vp, u, u, sz])
        _sig(L, "rtcOccludedNM", None, [vp, ctxp, vp, u, u, sz])
        for w in (4, 8, 16):
            _sig(L, f"rtcIntersect{w}", None, [vp, vp, ctxp, vp])
            _sig(L, f"rtcOccluded{w}", None, [vp, vp, ctxp, vp])
        self.has_ext = hasattr(L, "rtcxGetLaunchCount")
        if self.has_ext:
            _sig(L, "rtcxSetDeviceStream", None, [vp, vp])
            _sig(L, "rtcxSynchronizeDevice", None, [vp])
            _sig(L, "rtcxGetDeviceOrdinal", C.c_int, [vp])
            _sig(L, "rtcxGetDeviceGpuCount", C.c_int, [vp])
            _sig(L, "rtcxGetSceneBuildStats", C.c_int, [vp, C.POINTER(BuildStats)])
            _sig(L, "rtcxGetSceneImage", vp, [vp, C.POINTER(sz)])
            _sig(L, "rtcxSetSceneImage", None, [vp, vp, sz])
            _sig(L, "rtcxCopySceneImage", None, [vp, vp, sz])
            _sig(L, "rtcxSaveSceneImage", C.c_int, [vp, C.c_char_p])
            _sig(L, "rtcxLoadSceneImage", C.c_int, [vp, C.c_char_p])
            _sig(L, "rtcxIntersect1MCounted", None, [vp, ctxp, vp, u, sz, C.POINTER(TraceCounters)])
            _sig(L, "rtcxOccluded1MCounted", None, [vp, ctxp, vp, u, sz, C.POINTER(TraceCounters)])
            _sig(L, "rtcxGetLaunchCount", C.c_ulonglong, [])
            _sig(L, "rtcxGetTransferBytes", None, [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)])

    # ---- small conveniences used by tests and bench ----
    def new_device(self, cfg=""):
        d = self.lib.rtcNewDevice(cfg.encode())
        if not d:
            raise RuntimeError(f"rtcNewDevice failed, error {self.lib.rtcGetDeviceError(None)} ({self.path})")
        return d

    def context(self, coherent=False, inst_id=RTC_INVALID_GEOMETRY_ID):
        c = IntersectContext()
        c.flags = RTC_INTERSECT_CONTEXT_FLAG_COHERENT if coherent else RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT
        c.filter = None
        c.instID[0] = inst_id
        return c

    def add_mesh(self, device, scene, vertices, triangles, keep=None):
        """Attach a triangle mesh -- or, for an (m,4) index array, a quad mesh -- through shared buffers.  vertices (n,3)
        float32 (a 16-byte tail pad is added as the API demands), triangles (m,3) uint32.  Returns (geomID, geometry handle)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.uint32)
        vpad = np.zeros(v.size + 4, dtype=np.float32)
        vpad[:v.size] = v.ravel()
        quads = t.ndim == 2 and t.shape[1] == 4
        g = self.lib.rtcNewGeometry(device, RTC_GEOMETRY_TYPE_QUAD if quads else RTC_GEOMETRY_TYPE_TRIANGLE)
        self.lib.rtcSetSharedGeometryBuffer(g, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, vpad.ctypes.data, 0, 12, len(v))
        self.lib.rtcSetSharedGeometryBuffer(g, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT4 if quads else RTC_FORMAT_UINT3, t.ctypes.data, 0,
                                            16 if quads else 12, len(t))
        self.lib.rtcCommitGeometry(g)
        gid = self.lib.rtcAttachGeometry(scene, g)
        if keep is not None:
            keep.extend([vpad, t])                      # shared buffers must outlive the scene
        return gid, g

    def build_scene(self, device, meshes, flags=RTC_SCENE_FLAG_NONE):
        """meshes: list of (vertices, triangles).  Returns (scene, keepalive list)."""
        keep = []
        sc = self.lib.rtcNewScene(device)
        if flags:
            self.lib.rtcSetSceneFlags(sc, flags)
        for v, t in meshes:
            _, g = self.add_mesh(device, sc, v, t, keep)
            self.lib.rtcReleaseGeometry(g)
        self.lib.rtcCommitScene(sc)
        return sc, keep

    def add_instance(self, device, scene, instanced_scene, l2w, fmt=RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR):
        """Attach an instance of `instanced_scene` (committed) to `scene`.  l2w: 12 floats, column major
        (vx, vy, vz, p) unless another RTC_FORMAT_* is given.  Returns (geomID, geometry handle)."""
        m = np.ascontiguousarray(l2w, dtype=np.float32).ravel()
        g = self.lib.rtcNewGeometry(device, RTC_GEOMETRY_TYPE_INSTANCE)
        self.lib.rtcSetGeometryInstancedScene(g, instanced_scene)
        self.lib.rtcSetGeometryTransform(g, 0, fmt, m.ctypes.data)
        self.lib.rtcCommitGeometry(g)
        gid = self.lib.rtcAttachGeometry(scene, g)
        return gid, g

    def build_instanced(self, device, objects, base_meshes, instances, flags=RTC_SCENE_FLAG_NONE):
        """objects: list of mesh lists (one instanced scene each); base_meshes: triangle meshes attached to the
        top-level scene first (geomIDs 0..); instances: list of (object index, l2w).  Returns (top scene,
        [object scenes], keepalive)."""
        keep, obj = [], []
        for meshes in objects:
            sc, k = self.build_scene(device, meshes, flags)
            obj.append(sc)
            keep.append(k)
        top = self.lib.rtcNewScene(device)
        if flags:
            self.lib.rtcSetSceneFlags(top, flags)
        for v, t in base_meshes:
            _, g = self.add_mesh(device, top, v, t, keep)
            self.lib.rtcReleaseGeometry(g)
        for oi, m in instances:
            _, g = self.add_instance(device, top, obj[oi], m)
            self.lib.rtcReleaseGeometry(g)
        self.lib.rtcCommitScene(top)
        return top, obj, keep

    @staticmethod
    def _ptr(rays):
        if isinstance(rays, np.ndarray):
            return rays.ctypes.data, rays.strides[0], len(rays)
        raise TypeError("expected a numpy structured array; pass device pointers to *_ptr methods")

    def intersect(self, scene, rays, coherent=False, inst_id=RTC_INVALID_GEOMETRY_ID):
        p, stride, n = self._ptr(rays)
        ctx = self.context(coherent, inst_id)
        self.lib.rtcIntersect1M(scene, C.byref(ctx), p, n, stride)

    def occluded(self, scene, rays, coherent=False):
        p, stride, n = self._ptr(rays)
        ctx = self.context(coherent)
        self.lib.rtcOccluded1M(scene, C.byref(ctx), p, n, stride)

    def intersect_ptr(self, scene, ptr, n, stride=80, coherent=False):
        ctx = self.context(coherent)
        self.lib.rtcIntersect1M(scene, C.byref(ctx), ptr, n, stride)

    def occluded_ptr(self, scene, ptr, n, stride=48, coherent=False):
        ctx = self.context(coherent)
        self.lib.rtcOccluded1M(scene, C.byref(ctx), ptr, n, stride)

    def transfer_bytes(self, device):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self.lib.rtcxGetTransferBytes(device, C.byref(a), C.byref(b))
        return a.value, b.value

    def build_stats(self, scene):
        s = BuildStats()
        if self.lib.rtcxGetSceneBuildStats(scene, C.byref(s)) != 0:
            raise RuntimeError("rtcxGetSceneBuildStats failed")
        return {k: getattr(s, k) for k, _ in BuildStats._fields_}

    def intersect_counted(self, scene, rays_or_ptr, n=None, stride=80, occluded=False):
        if isinstance(rays_or_ptr, np.ndarray):
            p, stride, n = self._ptr(rays_or_ptr)
        else:
            p = rays_or_ptr
        ctx = self.context()
        c = TraceCounters()
        fn = self.lib.rtcxOccluded1MCounted if occluded else self.lib.rtcxIntersect1MCounted
        fn(scene, C.byref(ctx), p, n, stride, C.byref(c))
        return {k: getattr(c, k) for k, _ in TraceCounters._fields_}
