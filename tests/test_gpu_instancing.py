"""GPU parity for single-level instancing through the C ABI: golden vectors of the real
reference library, the oracle on seeded inputs, and the API rules around instance geometries."""
import ctypes as C

import numpy as np
import pytest

import cases
import instancing

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
pytestmark = pytest.mark.gpu
INV = 0xFFFFFFFF


def _release(product, top, objs):
    product.lib.rtcReleaseScene(top)
    for o in objs:
        product.lib.rtcReleaseScene(o)


@pytest.mark.parametrize("name", list(instancing.CASES))
def test_instancing_matches_reference_golden(product, gpu_device, name):
    c = instancing.CASES[name]()
    g = instancing.load_golden(name)
    top, objs, keep = product.build_instanced(gpu_device, c["objects"], c["base"], c["instances"], c["flags"])
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    b = rt.Bounds()
    product.lib.rtcGetSceneBounds(top, C.byref(b))
    ours_b = np.array([b.lower_x, b.lower_y, b.lower_z, b.upper_x, b.upper_y, b.upper_z], dtype=np.float32)
    assert np.allclose(ours_b, g["bounds"], rtol=1e-6, atol=1e-6), (ours_b, g["bounds"])
    r = g["rays"].copy()
    product.intersect(top, r)
    res = parity.compare_closest(r, g["closest"])
    assert res["pass"] and res["hits_ours"] > 1000, str(res)
    miss = g["closest"]["geomID"] == INV
    assert np.array_equal(r[miss].view(np.uint8), g["closest"][miss].view(np.uint8))     # misses / inactive rays untouched
    s = g["shadow_in"].copy()
    product.occluded(top, s)
    ro = parity.compare_occluded(s, g["shadow_out"])
    assert ro["pass"], str(ro)
    # coherent hint and the single-ray entry point take the same path
    r2 = g["rays"][:2000].copy()
    product.intersect(top, r2, coherent=True)
    assert parity.compare_closest(r2, g["closest"][:2000])["pass"]
    one = g["rays"][100:101].copy()
    ctx = product.context()
    product.lib.rtcIntersect1(top, C.byref(ctx), one.ctypes.data)
    assert parity.compare_closest(one, g["closest"][100:101])["pass"]
    _release(product, top, objs)


def test_instancing_device_resident_and_strided_streams(product, gpu_device):
    import torch
    c = instancing.CASES["inst_forest"]()
    g = instancing.load_golden("inst_forest")
    top, objs, keep = product.build_instanced(gpu_device, c["objects"], c["base"], c["instances"], c["flags"])
    n = len(g["rays"])
    d = torch.from_numpy(g["rays"].view(np.uint8).reshape(n, 80).copy()).cuda()
    product.intersect_ptr(top, d.data_ptr(), n)
    back = d.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE)
    assert parity.compare_closest(back, g["closest"])["pass"]
    wide = np.zeros((n, 100), dtype=np.uint8)                     # 100-byte stride: 4-byte aligned only -> unaligned kernel variant
    wide[:, :80] = g["rays"].view(np.uint8).reshape(n, 80)
    dw = torch.from_numpy(wide).cuda()
    product.intersect_ptr(top, dw.data_ptr(), n, stride=100)
    bw = np.ascontiguousarray(dw.cpu().numpy()[:, :80]).reshape(-1).view(rt.RAYHIT_DTYPE)
    assert np.array_equal(bw, back)
    _release(product, top, objs)


def test_instancing_against_oracle_many_instances(product, gpu_device, oracle):
    """A few hundred instances of a 32K-triangle object over a ground plane, incoherent + shadow rays."""
    obj = [fx.triangle_sphere((0.0, 0.0, 0.0), 0.45, 91)]
    base = [fx.displaced_plane(64, extent=12.0)]
    inst = []
    for x in range(-8, 9):
        for z in range(-8, 9):
            inst.append((0, instancing._xfm(0.37 * x + 0.11 * z, 0.05 * z, [1.0 + 0.03 * x, 0.8 + 0.02 * (x + z + 16), 1.0], [1.3 * x, 1.0 + 0.1 * ((x * z) % 3), 1.3 * z])))
    c = dict(objects=[obj], base=base, instances=inst, flags=0)
    top, objs, keep = product.build_instanced(gpu_device, c["objects"], c["base"], c["instances"], 0)
    st = product.build_stats(top)
    assert st["numPrimsValid"] == fx.num_tris(base) + len(inst)
    otop, handles = instancing.build_oracle(oracle, c)
    rays = np.concatenate([fx.incoherent_rays(20000, org=(0.3, 4.0, 0.2), seed=13),
                           fx.primary_rays(128, 128, org=(0.1, 25.0, 0.3), look=(0, -1, 0), up=(0, 0, 1))])
    a, w = rays.copy(), rays.copy()
    product.intersect(top, a)
    oracle.top_intersect(otop, w)
    res = parity.compare_closest(a, w)
    assert res["pass"] and res["hits_ours"] > 20000, str(res)
    hit = a["geomID"] != INV
    assert (a["instID"][hit] != INV).sum() > 5000
    sa = fx.shadow_rays(w)
    sw = sa.copy()
    product.occluded(top, sa)
    oracle.top_occluded(otop, sw)
    assert parity.compare_occluded(sa, sw)["pass"]
    oracle.free_top(otop)
    for h in handles:
        oracle.free(h)
    _release(product, top, objs)


def test_instance_api_rules_and_updates(product, gpu_device, oracle):
    L = product.lib
    obj_meshes = [fx.triangle_sphere((0.0, 0.0, 0.0), 0.5, 12)]
    obj, keep = product.build_scene(gpu_device, obj_meshes)
    top = L.rtcNewScene(gpu_device)
    m0 = instancing._xfm(0.2, 0.1, [1, 1, 1], [0.0, 0.0, 3.0])
    gid, g = product.add_instance(gpu_device, top, obj, m0)
    assert gid == 0
    L.rtcCommitScene(top)
    assert L.rtcGetDeviceError(gpu_device) == 0
    rays = fx.primary_rays(96, 96, org=(0, 0, -2), look=(0, 0, 1), up=(0, 1, 0))
    oh = oracle.build(obj_meshes)
    for m in (m0, instancing._xfm(1.1, -0.4, [1.5, 0.7, 1.2], [0.4, -0.2, 2.5])):
        L.rtcSetGeometryTransform(g, 0, rt.RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, m.ctypes.data)
        L.rtcCommitGeometry(g)
        L.rtcCommitScene(top)                                      # a moved instance: top-level rebuild
        ot = oracle.build_top(None, [(oh, m, 0)])
        a, w = rays.copy(), rays.copy()
        product.intersect(top, a, inst_id=7)                       # context instID is overwritten by instance hits only
        oracle.top_intersect(ot, w, 7)
        res = parity.compare_closest(a, w)
        assert res["pass"] and res["hits_ours"] > 100, str(res)
        assert (a["instID"][a["geomID"] != INV] == 0).all()
        oracle.free_top(ot)
    # the image of an instanced scene cannot be exported (it points into other scenes)
    n = C.c_size_t()
    assert not L.rtcxGetSceneImage(top, C.byref(n))
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    # multi-level instancing is rejected (RTC_MAX_INSTANCE_LEVEL_COUNT = 1)
    top2 = L.rtcNewScene(gpu_device)
    _, g2 = product.add_instance(gpu_device, top2, top, m0)
    L.rtcCommitScene(top2)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    # an uncommitted instanced scene is an error
    fresh = L.rtcNewScene(gpu_device)
    top3 = L.rtcNewScene(gpu_device)
    _, g3 = product.add_instance(gpu_device, top3, fresh, m0)
    L.rtcCommitScene(top3)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    # disabling the instance empties the scene
    L.rtcDisableGeometry(g)
    L.rtcCommitScene(top)
    a = rays.copy()
    product.intersect(top, a)
    assert np.array_equal(a, rays) and L.rtcGetDeviceError(gpu_device) == 0
    for h in (g, g2, g3):
        L.rtcReleaseGeometry(h)
    for s in (top, top2, top3, fresh, obj):
        L.rtcReleaseScene(s)
    oracle.free(oh)


def test_recommitted_instanced_scene_needs_top_commit(product, gpu_device, oracle):
    """The instance table holds device pointers into the instanced scene's image: after that scene is
    re-committed, queries on the top-level scene fail cleanly until it is committed again."""
    L = product.lib
    keep = []
    obj = L.rtcNewScene(gpu_device)
    v, t = fx.triangle_sphere((0.0, 0.0, 0.0), 0.5, 12)
    _, gm = product.add_mesh(gpu_device, obj, v, t, keep)
    L.rtcCommitScene(obj)
    top = L.rtcNewScene(gpu_device)
    m = instancing._xfm(0.0, 0.0, [1, 1, 1], [0.0, 0.0, 3.0])
    _, gi = product.add_instance(gpu_device, top, obj, m)
    L.rtcCommitScene(top)
    rays = fx.primary_rays(64, 64, org=(0, 0, -2), look=(0, 0, 1), up=(0, 1, 0))
    a = rays.copy()
    product.intersect(top, a)
    n0 = int((a["geomID"] != INV).sum())
    assert n0 > 100
    keep[0][:v.size] *= 2.0                                        # the sphere doubles in size
    L.rtcUpdateGeometryBuffer(gm, rt.RTC_BUFFER_TYPE_VERTEX, 0)
    L.rtcCommitGeometry(gm)
    L.rtcCommitScene(obj)
    b = rays.copy()
    product.intersect(top, b)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION and np.array_equal(b, rays)
    L.rtcCommitScene(top)
    product.intersect(top, b)
    assert L.rtcGetDeviceError(gpu_device) == 0 and int((b["geomID"] != INV).sum()) > 2 * n0
    oh = oracle.build([(keep[0][:v.size].reshape(-1, 3), t)])
    ot = oracle.build_top(None, [(oh, m, 0)])
    w = rays.copy()
    oracle.top_intersect(ot, w)
    assert parity.compare_closest(b, w)["pass"]
    oracle.free_top(ot)
    oracle.free(oh)
    L.rtcReleaseGeometry(gm)
    L.rtcReleaseGeometry(gi)
    L.rtcReleaseScene(top)
    L.rtcReleaseScene(obj)
