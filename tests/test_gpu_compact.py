"""RTC_SCENE_FLAG_COMPACT (reference: Triangle4i leaves, kernels/geometry/trianglei.h, chosen by
Scene::createTriangleAccel, scene.cpp:127-128): indexed 16-byte triangle records + a vertex pool inside the image.
The tree is the one the default layout gets, only the leaf storage differs, so every answer must be BIT-identical to the
default scene's -- and therefore within the north-star tolerances of the reference goldens."""
import ctypes as C

import numpy as np
import pytest

import cases

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
pytestmark = pytest.mark.gpu
COMPACT = rt.RTC_SCENE_FLAG_COMPACT


@pytest.mark.parametrize("name", ["sphere_small", "two_geoms", "garbage_prims", "robust_far_sphere", "edge_rays", "overlapping"])
def test_compact_scene_answers_like_the_default_layout(product, gpu_device, name):
    g = cases.load_golden(name)
    sc0, k0 = product.build_scene(gpu_device, g["meshes"], g["flags"])
    sc1, k1 = product.build_scene(gpu_device, g["meshes"], g["flags"] | COMPACT)
    s0, s1 = product.build_stats(sc0), product.build_stats(sc1)
    assert s1["numNodes"] == s0["numNodes"] and s1["numTris"] == s0["numTris"] and s1["sah"] == s0["sah"]
    assert s1["bytes"] < s0["bytes"]                                     # the point of the flag
    a, b = g["rays"].copy(), g["rays"].copy()
    product.intersect(sc0, a)
    product.intersect(sc1, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    if name != "overlapping":
        assert parity.compare_closest(b, g["closest"])["pass"]
    o0 = cases.occluded_by_group(lambda part: product.occluded(sc0, part), fx.to_ray(g["rays"]), g.get("groups"))
    o1 = cases.occluded_by_group(lambda part: product.occluded(sc1, part), fx.to_ray(g["rays"]), g.get("groups"))
    assert np.array_equal(o0.view(np.uint8), o1.view(np.uint8))
    assert parity.compare_occluded(o1, g["occl_self_out"])["disagree"] == 0
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    product.lib.rtcReleaseScene(sc0); product.lib.rtcReleaseScene(sc1)


def test_compact_quads_streams_image_and_refit(product, gpu_device):
    """Quads (second half flagged per triangle), device-resident and host-staged streams beyond the staging thresholds, image
    export / adoption (validated), structural check of the exported image, and refit through the vertex pool."""
    import torch
    from oracle import rq_image
    import quads
    meshes = [fx.displaced_plane(60, extent=3.0), fx.quad_plane((-2, 1.0, -2), (4, 0, 0), (0, 0.5, 4), 9, 7), fx.triangle_sphere((0, 1.5, 0), 0.6, 14)]
    L = product.lib
    scenes = []
    for flags in (0, COMPACT):
        sc = L.rtcNewScene(gpu_device)
        L.rtcSetSceneFlags(sc, flags)
        keep, geoms = [], []
        for v, t in meshes:
            _, gh = product.add_mesh(gpu_device, sc, v, t, keep)
            L.rtcSetGeometryBuildQuality(gh, rt.RTC_BUILD_QUALITY_REFIT); L.rtcCommitGeometry(gh)
            geoms.append(gh)
        L.rtcCommitScene(sc)
        scenes.append((sc, keep, geoms))
    (sc0, k0, g0), (sc1, k1, g1) = scenes
    img = rq_image.fetch(product, sc1)
    assert img.compact and img.check_structure()
    assert abs(img.sah()[0] - product.build_stats(sc1)["sah"]) <= 1e-6 * img.sah()[0]
    rays = np.tile(fx.incoherent_rays(50000, org=(0.1, 2.5, 0.2), seed=5), 6)        # 300 K rays: the pageable staging path engages
    a, b = rays.copy(), rays.copy()
    product.intersect(sc0, a); product.intersect(sc1, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and (a["geomID"] == 1).any()   # quad hits included
    d = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), 80).copy()).cuda()
    product.intersect_ptr(sc1, d.data_ptr(), len(rays))
    assert np.array_equal(d.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE), a)
    sh = fx.shadow_rays(a)
    s0, s1 = sh.copy(), sh.copy()
    product.occluded(sc0, s0); product.occluded(sc1, s1)
    assert np.array_equal(s0.view(np.uint8), s1.view(np.uint8))
    # image adoption: a byte copy of the compact image answers identically
    n = C.c_size_t(0)
    p = L.rtcxGetSceneImage(sc1, C.byref(n))
    sc2 = L.rtcNewScene(gpu_device)
    L.rtcxSetSceneImage(sc2, p, n.value)
    assert L.rtcGetDeviceError(gpu_device) == 0
    c = rays.copy(); product.intersect(sc2, c)
    assert np.array_equal(c.view(np.uint8), a.view(np.uint8))
    # refit: move the terrain, both layouts refit (no rebuild) and still agree bit for bit
    for (sc, keep, geoms) in scenes:
        keep[0][:meshes[0][0].size] += np.float32(0.05)
        L.rtcUpdateGeometryBuffer(geoms[0], rt.RTC_BUFFER_TYPE_VERTEX, 0); L.rtcCommitGeometry(geoms[0]); L.rtcCommitScene(sc)
        assert product.build_stats(sc)["refitCount"] == 1
    a, b = rays.copy(), rays.copy()
    product.intersect(sc0, a); product.intersect(sc1, b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert L.rtcGetDeviceError(gpu_device) == 0
    for sc in (sc0, sc1, sc2):
        L.rtcReleaseScene(sc)
    for _, _, geoms in scenes:
        for gh in geoms:
            L.rtcReleaseGeometry(gh)
