"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the headers
declare, the struct layouts are the reference's, and the host-side object model follows the
reference's error conventions and state machine (no compute without a GPU).
Reference behaviour being mirrored: kernels/common/rtcore.cpp, scene.cpp:595-653, geometry.cpp:80-107,
scene_triangle_mesh.cpp:35-80, device.cpp:258-316."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases

rt = cases.rt
ROOT = cases.ROOT


def declared_symbols():
    names = set()
    for h in ("include/embree3/rtcore.h", "include/rq_b200.h"):
        src = open(os.path.join(ROOT, h)).read()
        names |= set(re.findall(r"RTC_API[^;(]*?\b(rtcx?[A-Z]\w*)\s*\(", src))
    for w in (4, 8, 16):
        names |= {f"rtcIntersect{w}", f"rtcOccluded{w}"}
    return names


def test_library_exports_every_declared_symbol(product):
    out = subprocess.check_output(["nm", "-D", "--defined-only", product.path]).decode()
    exported = set(re.findall(r" T (\w+)", out))
    decl = declared_symbols()
    assert len(decl) > 70
    assert not (decl - exported), sorted(decl - exported)
    assert all(s.startswith("rtc") for s in exported), [s for s in exported if not s.startswith("rtc")]
    # out-of-scope entry points still link
    for s in ("rtcSetGeometryTransform", "rtcPointQuery", "rtcCollide", "rtcInterpolate", "rtcNewBVH", "rtcSetGeometryInstancedScene"):
        assert s in exported, s


def test_struct_layouts_match_reference_headers(tmp_path):
    """sizeof/offsetof of the wire structs (rtcore_ray.h:11-49: 48 + 32 bytes, 16-byte aligned)."""
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "embree3/rtcore.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct RTCRay), sizeof(struct RTCHit), sizeof(struct RTCRayHit),
         offsetof(struct RTCRay, tfar), offsetof(struct RTCRayHit, hit), offsetof(struct RTCHit, primID),
         sizeof(struct RTCRayHit4), sizeof(struct RTCRayHit8), sizeof(struct RTCRayHit16), sizeof(struct RTCBounds),
         sizeof(struct RTCIntersectContext), _Alignof(struct RTCRay8), _Alignof(struct RTCRay16));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = list(map(int, subprocess.check_output([str(exe)]).split()))
    assert vals == [48, 32, 80, 32, 48, 20, 320, 640, 1280, 32, 24, 32, 64], vals
    assert rt.RAY_DTYPE.itemsize == 48 and rt.RAYHIT_DTYPE.itemsize == 80


def test_header_compiles_as_cpp_and_forwarders(tmp_path):
    src = tmp_path / "h.cpp"
    src.write_text('#include "embree3/rtcore_ray.h"\n#include "embree3/rtcore_scene.h"\n#include "embree3/rtcore_geometry.h"\n'
                   '#include "rq_b200.h"\nint main(){ RTCIntersectContext c; rtcInitIntersectContext(&c); RTCRayHit r; (void)r; return c.instID[0]==RTC_INVALID_GEOMETRY_ID?0:1; }\n')
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_no_gpu_means_no_device(product):
    """Without CUDA rtcNewDevice fails loudly (no CPU fallback); with CUDA it succeeds."""
    import torch
    d = product.lib.rtcNewDevice(b"")
    if torch.cuda.is_available():
        assert d
        product.lib.rtcReleaseDevice(d)
    else:
        assert not d
        assert product.lib.rtcGetDeviceError(None) == rt.RTC_ERROR_UNKNOWN
        assert product.lib.rtcGetDeviceError(None) == rt.RTC_ERROR_NONE      # reading clears (device.cpp:273-286)


@pytest.fixture()
def hostdev(product):
    """A device object without GPU requirements: host-side API logic only (commit/query would fail)."""
    d = product.lib.rtcNewDevice(b"allow_no_gpu=1")
    assert d
    yield d
    product.lib.rtcReleaseDevice(d)


def err(product, d):
    return product.lib.rtcGetDeviceError(d)


def test_error_slot_first_error_wins_and_callback(product, hostdev):
    L = product.lib
    seen = []
    CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p)
    cb = CB(lambda p, code, msg: seen.append((code, msg.decode())))
    L.rtcSetDeviceErrorFunction.argtypes = [C.c_void_p, CB, C.c_void_p]
    L.rtcSetDeviceErrorFunction(hostdev, cb, None)
    assert not L.rtcNewGeometry(hostdev, 2)                                   # RTC_GEOMETRY_TYPE_GRID: unsupported type -> INVALID_OPERATION
    L.rtcCommitGeometry(None)                                                 # NULL handle -> INVALID_ARGUMENT, thread slot
    g = L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    L.rtcSetGeometryTimeStepCount(g, 2)                                       # second error on the device: not recorded
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    assert err(product, hostdev) == rt.RTC_ERROR_NONE
    assert err(product, None) == rt.RTC_ERROR_INVALID_ARGUMENT
    assert [c for c, _ in seen] == [rt.RTC_ERROR_INVALID_OPERATION, rt.RTC_ERROR_INVALID_OPERATION]
    L.rtcSetDeviceErrorFunction(hostdev, CB(0), None)
    L.rtcReleaseGeometry(g)


def test_buffer_rules(product, hostdev):
    """scene_triangle_mesh.cpp:35-80: alignment, formats, slots."""
    L = product.lib
    g = L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    v = np.zeros(16, np.float32)
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data + 2, 0, 12, 3)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION           # not 4-byte aligned
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 10, 3)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION           # stride not a multiple of 4
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_UINT3, v.ctypes.data, 0, 12, 3)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION           # wrong vertex format
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 1, rt.RTC_FORMAT_UINT3, v.ctypes.data, 0, 12, 1)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_ARGUMENT            # index slot != 0
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 12, 1)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION           # wrong index format
    L.rtcSetSharedGeometryBuffer(g, 99, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 12, 1)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_ARGUMENT            # unknown buffer type
    p = L.rtcSetNewGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, 12, 5)
    assert p and p % 16 == 0 and L.rtcGetGeometryBufferData(g, rt.RTC_BUFFER_TYPE_VERTEX, 0) == p
    assert err(product, hostdev) == rt.RTC_ERROR_NONE
    b = L.rtcNewBuffer(hostdev, 256)
    assert L.rtcGetBufferData(b)
    L.rtcSetGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, b, 4, 12, 8)
    assert L.rtcGetGeometryBufferData(g, rt.RTC_BUFFER_TYPE_INDEX, 0) == L.rtcGetBufferData(b) + 4
    L.rtcReleaseBuffer(b)
    L.rtcReleaseGeometry(g)
    assert err(product, hostdev) == rt.RTC_ERROR_NONE


def test_geometry_ids_lowest_free_first(product, hostdev):
    """scene.cpp:595-623 + common/sys/alloc.h:100-125: smallest previously freed ID first."""
    L = product.lib
    sc = L.rtcNewScene(hostdev)
    gs = [L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_TRIANGLE) for _ in range(6)]
    ids = [L.rtcAttachGeometry(sc, g) for g in gs[:4]]
    assert ids == [0, 1, 2, 3]
    L.rtcDetachGeometry(sc, 2)
    L.rtcDetachGeometry(sc, 0)
    assert L.rtcAttachGeometry(sc, gs[4]) == 0
    assert L.rtcAttachGeometry(sc, gs[5]) == 2
    assert L.rtcGetGeometry(sc, 2) == gs[5] and L.rtcGetGeometry(sc, 1) == gs[1]
    L.rtcAttachGeometryByID(sc, gs[0], 10)
    assert L.rtcGetGeometry(sc, 10) == gs[0]
    assert L.rtcAttachGeometry(sc, gs[2]) == 4                               # ids 4..9 are free now, lowest first
    L.rtcAttachGeometryByID(sc, gs[3], 1)                                    # occupied
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcDetachGeometry(sc, 77)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    for g in gs:
        L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)
    assert err(product, hostdev) == rt.RTC_ERROR_NONE


def test_commit_state_machine_without_gpu(product, hostdev):
    L = product.lib
    sc = L.rtcNewScene(hostdev)
    ctx = product.context()
    r = rt.new_rays(4)
    L.rtcIntersect1M(sc, C.byref(ctx), r.ctypes.data, 4, 80)                 # scene.cpp:13,30 missing_rtcCommit
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    b = rt.Bounds()
    L.rtcGetSceneBounds(sc, C.byref(b))
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    g = L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    L.rtcAttachGeometry(sc, g)
    L.rtcCommitScene(sc)                                                     # geometry.cpp:103-107
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcCommitGeometry(g)
    import torch
    if not torch.cuda.is_available():
        L.rtcCommitScene(sc)                                                 # no device: loud failure, not a CPU build
        assert err(product, hostdev) == rt.RTC_ERROR_UNKNOWN
    L.rtcSetSceneFlags(sc, rt.RTC_SCENE_FLAG_ROBUST)
    assert L.rtcGetSceneFlags(sc) == rt.RTC_SCENE_FLAG_ROBUST
    L.rtcSetSceneBuildQuality(sc, 7)
    assert err(product, hostdev) == rt.RTC_ERROR_UNKNOWN                      # the reference throws a plain runtime_error (rtcore.cpp:233)
    L.rtcSetGeometryTessellationRate.argtypes = [C.c_void_p, C.c_float]
    L.rtcSetGeometryTessellationRate(g, 4.0)                                 # out-of-scope entry point
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    assert L.rtcGetDeviceProperty(hostdev, 0) == 31201 and L.rtcGetDeviceProperty(hostdev, 96) == 1
    assert L.rtcGetDeviceProperty(hostdev, 66) == 0 and L.rtcGetDeviceProperty(hostdev, 35) == 1
    assert L.rtcGetDeviceProperty(hostdev, 97) == 1 and L.rtcGetDeviceProperty(hostdev, 98) == 0   # quads yes, subdivision no
    L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)


def test_oracle_is_not_on_the_product_path():
    """Nothing under the package may reference oracle/ (the checker is never the thing shipped)."""
    pk = os.path.join(ROOT, "embree-aarch64_b200")
    for dp, _, fs in os.walk(pk):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "rq_oracle" not in txt and "oracle/" not in txt and "libembree3_ref" not in txt, os.path.join(dp, f)
    out = subprocess.check_output(["ldd", os.path.join(pk, "lib", "libembree3.so")]).decode()
    assert "oracle" not in out and "torch" not in out


def test_instance_geometry_host_logic(product, hostdev):
    """rtcSetGeometryTransform / rtcGetGeometryTransform formats (rtcore.cpp:1008-1069) and instance-only calls."""
    _instance_host_logic(product, hostdev)


def test_instance_geometry_host_logic_is_the_references(reflib):
    """The same call sequence against the real reference library: the expectations above are its behaviour, not ours."""
    d = reflib.new_device("")
    _instance_host_logic(reflib, d)
    reflib.lib.rtcReleaseDevice(d)


def _instance_host_logic(product, hostdev):
    L = product.lib
    g = L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_INSTANCE)
    assert g and err(product, hostdev) == 0
    col = np.arange(1, 13, dtype=np.float32)                        # vx=(1,2,3) vy=(4,5,6) vz=(7,8,9) p=(10,11,12)
    L.rtcSetGeometryTransform(g, 0, rt.RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, col.ctypes.data)
    out = np.zeros(16, dtype=np.float32)
    L.rtcGetGeometryTransform(g, 0.0, rt.RTC_FORMAT_FLOAT3X4_ROW_MAJOR, out.ctypes.data)
    assert np.array_equal(out[:12], np.array([1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12], dtype=np.float32))
    L.rtcGetGeometryTransform(g, 0.0, rt.RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR, out.ctypes.data)
    assert np.array_equal(out, np.array([1, 2, 3, 0, 4, 5, 6, 0, 7, 8, 9, 0, 10, 11, 12, 1], dtype=np.float32))
    row = np.array([1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12], dtype=np.float32)
    L.rtcSetGeometryTransform(g, 0, rt.RTC_FORMAT_FLOAT3X4_ROW_MAJOR, row.ctypes.data)
    L.rtcGetGeometryTransform(g, 0.0, rt.RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, out.ctypes.data)
    assert np.array_equal(out[:12], col)
    m44 = np.array([1, 2, 3, 0, 4, 5, 6, 0, 7, 8, 9, 0, 10, 11, 12, 1], dtype=np.float32)
    L.rtcSetGeometryTransform(g, 0, rt.RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR, m44.ctypes.data)
    L.rtcGetGeometryTransform(g, 0.0, rt.RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, out.ctypes.data)
    assert np.array_equal(out[:12], col) and err(product, hostdev) == 0
    L.rtcSetGeometryTransform(g, 0, 0x9133, col.ctypes.data)                     # FLOAT3X3: invalid matrix format
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetGeometryTransform(g, 1, rt.RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, col.ctypes.data)   # time step 1: no motion blur
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    t = L.rtcNewGeometry(hostdev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    sc = L.rtcNewScene(hostdev)
    L.rtcSetGeometryInstancedScene(t, sc)                                         # not an instance geometry
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetGeometryInstancedScene(g, None)
    assert err(product, hostdev) == rt.RTC_ERROR_INVALID_ARGUMENT
    L.rtcSetGeometryInstancedScene(g, sc)
    L.rtcSetGeometryInstancedScene(g, sc)                                         # replacing keeps the reference counts balanced
    assert err(product, hostdev) == 0
    L.rtcReleaseScene(sc)                                                         # the geometry still holds the instanced scene
    L.rtcReleaseGeometry(g)
    L.rtcReleaseGeometry(t)


def _buffer_format_rules(lib, dev):
    """scene_triangle_mesh.cpp:35-80 / scene_quad_mesh.cpp:35-80: index format per geometry type, vertex format, alignment."""
    L = lib.lib
    v = np.zeros(64, dtype=np.float32)
    i = np.zeros(64, dtype=np.uint32)
    tri = L.rtcNewGeometry(dev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    quad = L.rtcNewGeometry(dev, rt.RTC_GEOMETRY_TYPE_QUAD)
    assert tri and quad and L.rtcGetDeviceError(dev) == 0
    L.rtcSetSharedGeometryBuffer(tri, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, i.ctypes.data, 0, 12, 4)
    L.rtcSetSharedGeometryBuffer(quad, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT4, i.ctypes.data, 0, 16, 4)
    L.rtcSetSharedGeometryBuffer(quad, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 12, 4)
    assert L.rtcGetDeviceError(dev) == 0
    L.rtcSetSharedGeometryBuffer(tri, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT4, i.ctypes.data, 0, 16, 4)
    assert L.rtcGetDeviceError(dev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetSharedGeometryBuffer(quad, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, i.ctypes.data, 0, 12, 4)
    assert L.rtcGetDeviceError(dev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetSharedGeometryBuffer(quad, rt.RTC_BUFFER_TYPE_VERTEX, 0, 0x9002, v.ctypes.data, 0, 8, 4)          # RTC_FORMAT_FLOAT2
    assert L.rtcGetDeviceError(dev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetSharedGeometryBuffer(tri, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 14, 4)   # stride not a multiple of 4
    assert L.rtcGetDeviceError(dev) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcSetSharedGeometryBuffer(tri, rt.RTC_BUFFER_TYPE_VERTEX, 1, rt.RTC_FORMAT_FLOAT3, v.ctypes.data, 0, 12, 4)   # slot 1 without time steps
    assert L.rtcGetDeviceError(dev) != 0
    L.rtcReleaseGeometry(tri)
    L.rtcReleaseGeometry(quad)


def test_buffer_format_rules(product, hostdev):
    _buffer_format_rules(product, hostdev)


def test_buffer_format_rules_are_the_references(reflib):
    d = reflib.new_device("")
    _buffer_format_rules(reflib, d)
    reflib.lib.rtcReleaseDevice(d)


def _scene_state_rules(lib, dev):
    """Geometry-ID allocation (lowest free first, scene.cpp:595-623), attach / detach / get errors, quality and time-step
    arguments, queries on an uncommitted scene, NULL handles -- one call sequence, the observations returned as a list."""
    L = lib.lib
    out = []

    def e():
        return L.rtcGetDeviceError(dev)
    sc = L.rtcNewScene(dev)
    gs = [L.rtcNewGeometry(dev, rt.RTC_GEOMETRY_TYPE_TRIANGLE) for _ in range(5)]
    out.append(("attach0", L.rtcAttachGeometry(sc, gs[0]), e()))
    out.append(("attach1", L.rtcAttachGeometry(sc, gs[1]), e()))
    L.rtcDetachGeometry(sc, 0); out.append(("detach0", e()))
    out.append(("attach2 takes the lowest free id", L.rtcAttachGeometry(sc, gs[2]), e()))
    L.rtcAttachGeometryByID(sc, gs[3], 5); out.append(("by id 5", e()))
    out.append(("attach4", L.rtcAttachGeometry(sc, gs[4]), e()))
    L.rtcAttachGeometryByID(sc, gs[0], 5); out.append(("by id 5 again", e()))
    L.rtcDetachGeometry(sc, 3); out.append(("detach unused id", e()))
    L.rtcDetachGeometry(sc, 77); out.append(("detach out of range", e()))
    out.append(("get 1", bool(L.rtcGetGeometry(sc, 1)), e()))
    out.append(("get unused id", bool(L.rtcGetGeometry(sc, 3)), e()))
    L.rtcSetSceneBuildQuality(sc, 3); out.append(("scene quality REFIT is invalid", e()))
    L.rtcSetGeometryBuildQuality(gs[1], 3); out.append(("geometry quality REFIT", e()))
    L.rtcSetGeometryBuildQuality(gs[1], 9); out.append(("geometry quality 9", e()))
    L.rtcSetGeometryTimeStepCount(gs[1], 0); out.append(("time steps 0", e()))
    L.rtcSetGeometryTimeStepCount(gs[1], 1); out.append(("time steps 1", e()))
    L.rtcSetGeometryTimeStepCount(gs[1], 1000); out.append(("time steps 1000", e()))
    r = rt.new_rays(4)
    ctx = lib.context()
    L.rtcIntersect1M(sc, C.byref(ctx), r.ctypes.data, 4, 80); out.append(("query on an uncommitted scene", e()))
    b = rt.Bounds()
    L.rtcGetSceneBounds(sc, C.byref(b)); out.append(("bounds of an uncommitted scene", e()))
    L.rtcCommitScene(None); out.append(("commit NULL", e(), L.rtcGetDeviceError(None)))
    for g in gs:
        L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)
    return out


def test_scene_state_rules_equal_the_references(product, hostdev, reflib):
    d = reflib.new_device("")
    want = _scene_state_rules(reflib, d)
    reflib.lib.rtcReleaseDevice(d)
    got = _scene_state_rules(product, hostdev)
    assert got == want, [(a, b) for a, b in zip(got, want) if a != b]
    assert ("attach2 takes the lowest free id", 0, 0) in got and ("by id 5 again", rt.RTC_ERROR_INVALID_OPERATION) in got


def _buffer_and_property_rules(lib, dev):
    """Buffers, geometry buffers (alignment, slots, unknown types), user data, masks, device properties, scene flags, NULL handles:
    one call sequence, the observations returned as a list."""
    L = lib.lib
    out = []
    def e(): return L.rtcGetDeviceError(dev)
    L.rtcGetGeometryUserData.restype = C.c_void_p; L.rtcGetGeometryUserData.argtypes=[C.c_void_p]
    L.rtcSetGeometryUserData.argtypes=[C.c_void_p, C.c_void_p]
    L.rtcSetGeometryMask.argtypes=[C.c_void_p, C.c_uint]
    L.rtcSetDeviceProperty.argtypes=[C.c_void_p, C.c_int, C.c_ssize_t]
    g = L.rtcNewGeometry(dev, rt.RTC_GEOMETRY_TYPE_TRIANGLE)
    out.append(("new shared buffer NULL", bool(L.rtcNewSharedBuffer(dev, None, 64)), e()))
    b = L.rtcNewBuffer(dev, 256); out.append(("new buffer", bool(b), e()))
    out.append(("buffer data", bool(L.rtcGetBufferData(b)), e()))
    L.rtcSetGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, b, 2, 12, 4); out.append(("unaligned offset", e()))
    L.rtcSetGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, b, 0, 12, 4); out.append(("vertex buffer ok", e()))
    L.rtcSetGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0, rt.RTC_FORMAT_FLOAT3, None, 0, 12, 4); out.append(("NULL buffer", e()))
    L.rtcSetGeometryBuffer(g, 77, 0, rt.RTC_FORMAT_FLOAT3, b, 0, 12, 4); out.append(("buffer type 77", e()))
    L.rtcSetGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 1, rt.RTC_FORMAT_UINT3, b, 0, 12, 4); out.append(("index slot 1", e()))
    out.append(("get vertex data", bool(L.rtcGetGeometryBufferData(g, rt.RTC_BUFFER_TYPE_VERTEX, 0)), e()))
    out.append(("get index data unset", bool(L.rtcGetGeometryBufferData(g, rt.RTC_BUFFER_TYPE_INDEX, 0)), e()))
    out.append(("get data type 77", bool(L.rtcGetGeometryBufferData(g, 77, 0)), e()))
    p = L.rtcSetNewGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, 12, 10); out.append(("set new index buffer", bool(p), e()))
    L.rtcUpdateGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0); out.append(("update vertex", e()))
    L.rtcUpdateGeometryBuffer(g, 77, 0); out.append(("update type 77", e()))
    L.rtcSetGeometryMask(g, 0xF0); out.append(("mask", e()))
    L.rtcSetGeometryUserData(g, 0x1234); out.append(("userdata", L.rtcGetGeometryUserData(g), e()))
    L.rtcEnableGeometry(g); L.rtcDisableGeometry(g); L.rtcDisableGeometry(g); out.append(("enable/disable", e()))
    L.rtcCommitGeometry(g); out.append(("commit geometry", e()))
    out.append(("prop 999", L.rtcGetDeviceProperty(dev, 999), e()))
    L.rtcSetDeviceProperty(dev, 999, 1); out.append(("set prop 999", e()))
    L.rtcSetDeviceProperty(dev, 0, 1); out.append(("set prop version", e()))
    sc = L.rtcNewScene(dev)
    out.append(("flags default", L.rtcGetSceneFlags(sc), e()))
    L.rtcSetSceneFlags(sc, 7); out.append(("flags 7", L.rtcGetSceneFlags(sc), e()))
    d2 = L.rtcGetSceneDevice(sc); out.append(("scene device", d2 == dev, e())); L.rtcReleaseDevice(d2)
    L.rtcAttachGeometry(None, g); out.append(("attach NULL scene", e(), L.rtcGetDeviceError(None)))
    L.rtcAttachGeometry(sc, None); out.append(("attach NULL geom", e(), L.rtcGetDeviceError(None)))
    L.rtcAttachGeometryByID(sc, g, 0xFFFFFFFF); out.append(("by id INVALID", e()))
    L.rtcReleaseBuffer(b); L.rtcReleaseGeometry(g); L.rtcReleaseScene(sc)
    return out


def test_buffer_and_property_rules_equal_the_references(product, hostdev, reflib):
    d = reflib.new_device("")
    want = _buffer_and_property_rules(reflib, d)
    reflib.lib.rtcReleaseDevice(d)
    got = _buffer_and_property_rules(product, hostdev)
    assert got == want, [(a, b) for a, b in zip(got, want) if a != b]
