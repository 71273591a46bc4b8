"""GPU parity for quad meshes through the C ABI: golden vectors of the real reference library, refit."""
import ctypes as C

import numpy as np
import pytest

import cases
import quads

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
pytestmark = pytest.mark.gpu
INV = 0xFFFFFFFF


@pytest.mark.parametrize("name", list(quads.CASES))
def test_quads_match_reference_golden(product, gpu_device, name):
    c = quads.CASES[name]()
    g = quads.load_golden(name)
    sc, keep = product.build_scene(gpu_device, c["meshes"], c["flags"])
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    st = product.build_stats(sc)
    assert st["numPrimsIn"] == fx.num_tris(c["meshes"]) and st["numPrimsValid"] < st["numPrimsIn"]      # dropped quads
    b = rt.Bounds()
    product.lib.rtcGetSceneBounds(sc, C.byref(b))
    ours_b = np.array([b.lower_x, b.lower_y, b.lower_z, b.upper_x, b.upper_y, b.upper_z], dtype=np.float32)
    assert np.array_equal(ours_b, g["bounds"]), (ours_b, g["bounds"])
    r = g["rays"].copy()
    product.intersect(sc, r)
    res = parity.compare_closest(r, g["closest"])
    assert res["pass"] and res["hits_ours"] > 5000, str(res)
    miss = g["closest"]["geomID"] == INV
    assert np.array_equal(r[miss].view(np.uint8), g["closest"][miss].view(np.uint8))
    s = g["shadow_in"].copy()
    product.occluded(sc, s)
    ro = parity.compare_occluded(s, g["shadow_out"])
    assert ro["pass"], str(ro)
    product.lib.rtcReleaseScene(sc)


def test_quad_mesh_refit_and_api(product, gpu_device, oracle):
    L = product.lib
    v0, q = quads.bumpy_quads(40, 3.0)
    sc = L.rtcNewScene(gpu_device)
    keep = []
    gid, g = product.add_mesh(gpu_device, sc, v0, q, keep)
    L.rtcSetGeometryBuildQuality(g, rt.RTC_BUILD_QUALITY_REFIT)
    L.rtcCommitGeometry(g)
    L.rtcCommitScene(sc)
    assert L.rtcGetDeviceError(gpu_device) == 0
    rays = np.concatenate([fx.incoherent_rays(20000, org=(0.1, 2.0, 0.2), seed=4),
                           fx.primary_rays(96, 96, org=(0.3, 6.0, 0.2), look=(0, -1, 0), up=(0, 0, 1))])
    for step in range(3):
        v = v0.copy()
        v[:, 1] += np.float32(0.3 * step) * np.sin(v0[:, 0] * 2 + step).astype(np.float32)
        if step == 2:
            v[55] = np.inf                                         # the quads around vertex 55 vanish
        if step:
            keep[0][:v.size] = v.ravel()
            L.rtcUpdateGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0)
            L.rtcCommitGeometry(g)
            L.rtcCommitScene(sc)
            assert product.build_stats(sc)["refitCount"] == step
        h = oracle.build([(v, q)])
        a, w = rays.copy(), rays.copy()
        product.intersect(sc, a)
        oracle.intersect(h, w)
        res = parity.compare_closest(a, w)
        assert res["pass"] and res["hits_ours"] > 10000, str(res)
        oracle.free(h)
    # a quad mesh wants RTC_FORMAT_UINT4 indices, a triangle mesh RTC_FORMAT_UINT3 (scene_quad_mesh.cpp / scene_triangle_mesh.cpp setBuffer)
    L.rtcSetSharedGeometryBuffer(g, rt.RTC_BUFFER_TYPE_INDEX, 0, rt.RTC_FORMAT_UINT3, keep[1].ctypes.data, 0, 12, 10)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)
