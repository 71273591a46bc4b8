"""RTC_BUILD_QUALITY_HIGH (reference: the spatial-split builders chosen by Scene::createTriangleAccel at HIGH
quality, scene.cpp:103-104 -> BVHNBuilderFastSpatialSAH, primrefgen_presplit.h, splitter.h): large triangles are pre-split into
several references with clipped boxes, treelets are 512 triangles and their last levels get an exact sweep SAH.  Answers must
not change; the tree must get better where long triangles would otherwise drag huge boxes through the hierarchy."""
import numpy as np
import pytest

import cases

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
pytestmark = pytest.mark.gpu
INV = 0xFFFFFFFF


def _scene_with_long_triangles():
    plane = fx.displaced_plane(96, extent=4.0)                                  # ~18 K triangles of ~0.08 units
    floor_v = np.array([[-30, -1, -30], [30, -1, -30], [30, -1, 30], [-30, -1, 30]], dtype=np.float32)
    floor_t = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)                 # two 60-unit axis-aligned triangles: large, but they fill their boxes
    # 240 long thin slivers crossing the terrain's volume in all diagonal directions: their boxes cover large parts of the scene
    rs = fx.RandomSampler(np.arange(240), 23)
    c = np.stack([rs.get_float() * 6 - 3, rs.get_float() * 1.2 - 0.4, rs.get_float() * 6 - 3], 1).astype(np.float32)
    d = np.stack([rs.get_float() * 2 - 1, rs.get_float() * 0.8 - 0.4, rs.get_float() * 2 - 1], 1).astype(np.float32)
    d = d / np.linalg.norm(d, axis=1, keepdims=True) * (2.0 + 2.0 * rs.get_float()[:, None])
    w = np.stack([rs.get_float() - 0.5, rs.get_float() - 0.5, rs.get_float() - 0.5], 1).astype(np.float32) * 0.04
    sl_v = np.stack([c - d, c + d, c + d + w], 1).reshape(-1, 3).astype(np.float32)
    sl_t = np.arange(3 * 240, dtype=np.uint32).reshape(-1, 3)
    return [plane, (floor_v, floor_t), (sl_v, sl_t)]


def _build(product, dev, meshes, quality):
    L = product.lib
    sc = L.rtcNewScene(dev)
    L.rtcSetSceneBuildQuality(sc, quality)
    keep = []
    for v, t in meshes:
        _, g = product.add_mesh(dev, sc, v, t, keep)
        L.rtcReleaseGeometry(g)
    L.rtcCommitScene(sc)
    assert L.rtcGetDeviceError(dev) == 0
    return sc, keep


def test_presplit_keeps_answers_and_tightens_the_tree(product, oracle):
    from oracle import rq_image
    dev = product.new_device("")
    meshes = _scene_with_long_triangles()
    med, k0 = _build(product, dev, meshes, rt.RTC_BUILD_QUALITY_MEDIUM)
    high, k1 = _build(product, dev, meshes, rt.RTC_BUILD_QUALITY_HIGH)
    s0, s1 = product.build_stats(med), product.build_stats(high)
    ntris = fx.num_tris(meshes)
    assert s0["numSplitRefs"] == 0 and s0["numTris"] == ntris
    assert s1["numSplitRefs"] >= 240 and s1["numTris"] == ntris + s1["numSplitRefs"]      # the slivers became many references ...
    assert s1["numSplitRefs"] <= 240 * 63                                                     # ... the axis-aligned floor triangles did not
    img = rq_image.fetch(product, high)
    keys = np.concatenate([(np.uint64(g) << np.uint64(33)) | (np.arange(len(t), dtype=np.uint64) << np.uint64(1)) for g, (v, t) in enumerate(meshes)])
    assert img.check_structure(keys, presplit=True)
    assert abs(img.sah()[0] - s1["sah"]) <= 1e-6 * s1["sah"]
    # rays from above (hit terrain, slivers or the floor next to the terrain) and from the side (graze the long triangles)
    r1 = fx.incoherent_rays(60000, org=(0.3, 3.0, -0.2), seed=13)
    rs = fx.RandomSampler(np.arange(60000), 17)
    o = np.stack([rs.get_float() * 50 - 25, rs.get_float() * 3 - 0.5, rs.get_float() * 50 - 25], 1).astype(np.float32)
    d = np.stack([rs.get_float() * 2 - 1, rs.get_float() * 0.6 - 0.5, rs.get_float() * 2 - 1], 1).astype(np.float32)
    r2 = fx._set(rt.new_rays(60000), o, d, 0.0, np.inf)
    rays = np.concatenate([r1, r2])
    a, b, w = rays.copy(), rays.copy(), rays.copy()
    product.intersect(med, a)
    product.intersect(high, b)
    h = oracle.build(meshes)
    oracle.intersect(h, w)
    oracle.free(h)
    assert parity.compare_closest(a, w)["pass"]
    res = parity.compare_closest(b, w)
    assert res["pass"], res
    assert (b["geomID"] == 1).sum() > 1000 and (b["geomID"] == 2).sum() > 100           # floor and slivers are hit
    sa, sb = fx.to_ray(rays), fx.to_ray(rays)
    product.occluded(med, sa); product.occluded(high, sb)
    assert parity.compare_occluded(sb, sa)["disagree"] == 0
    c0 = product.intersect_counted(med, rays.copy())
    c1 = product.intersect_counted(high, rays.copy())
    print("nodes/ray MEDIUM %.2f HIGH %.2f, tris/ray %.2f %.2f, SAH %.2f %.2f" % (c0["nodes"] / c0["rays"], c1["nodes"] / c1["rays"],
          c0["tris"] / c0["rays"], c1["tris"] / c1["rays"], s0["sah"], s1["sah"]))
    assert c1["nodes"] + c1["tris"] < c0["nodes"] + c0["tris"]                              # less traversal work per ray
    assert s1["sah"] < s0["sah"]
    product.lib.rtcReleaseScene(med); product.lib.rtcReleaseScene(high); product.lib.rtcReleaseDevice(dev)


def test_high_quality_on_uniform_tessellation_splits_nothing(product):
    dev = product.new_device("")
    meshes = fx.scene_c2(0.2)
    sc, keep = _build(product, dev, meshes, rt.RTC_BUILD_QUALITY_HIGH)
    st = product.build_stats(sc)
    assert st["numSplitRefs"] == 0 and st["numTris"] == fx.num_tris(meshes) and st["numTreelets"] > 0
    product.lib.rtcReleaseScene(sc); product.lib.rtcReleaseDevice(dev)
