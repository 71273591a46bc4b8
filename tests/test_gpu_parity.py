"""GPU parity tests: the CUDA path, called through the C ABI, against the golden vectors of the real
reference library, against the oracle on identical inputs, and -- at BASELINE.json's full sizes --
through size-independent properties.  Tolerances are the north star's: hit/miss agreement >= 99.99 %
(disagreements only within eps of an edge/vertex), geomID/primID exact on agreed hits, t within
1e-5 relative, u/v within 1e-4 (embree-aarch64_b200/parity.py)."""
import ctypes as C
import threading

import numpy as np
import pytest

import cases

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
pytestmark = pytest.mark.gpu
ULP = np.float32(1.1920928955078125e-07)
ALL = list(cases.CASES)
INV = 0xFFFFFFFF


def build(product, dev, g):
    return product.build_scene(dev, g["meshes"], g["flags"])


@pytest.mark.parametrize("name", ALL)
def test_closest_hit_matches_reference_golden(product, gpu_device, name):
    g = cases.load_golden(name)
    sc, keep = build(product, gpu_device, g)
    r = g["rays"].copy()
    product.intersect(sc, r)
    res = parity.compare_closest(r, g["closest"])
    if name == "overlapping":
        assert res["hitmiss_disagree"] == 0 and res["id_disagree_unexplained"] == 0 and res["t_out_of_tol"] == 0, res
    else:
        assert res["pass"], res
    miss = g["closest"]["geomID"] == INV
    assert np.array_equal(r[miss].view(np.uint8), g["closest"][miss].view(np.uint8))     # misses / inactive rays untouched
    for k in ("org_x", "org_y", "org_z", "tnear", "dir_x", "dir_y", "dir_z", "time", "mask", "id", "flags"):
        assert np.array_equal(r[k].view(np.uint32), g["rays"][k].view(np.uint32)), k     # inputs never modified
    product.lib.rtcReleaseScene(sc)


@pytest.mark.parametrize("name", ALL)
def test_occluded_matches_reference_golden(product, gpu_device, name):
    g = cases.load_golden(name)
    sc, keep = build(product, gpu_device, g)
    s = g["shadow_in"].copy()
    product.occluded(sc, s)
    assert parity.compare_occluded(s, g["shadow_out"])["disagree"] == 0
    s2 = cases.occluded_by_group(lambda part: product.occluded(sc, part), fx.to_ray(g["rays"]), g.get("groups"))
    res = parity.compare_occluded(s2, g["occl_self_out"])
    assert res["disagree"] == 0 and res["untouched_ok"], res
    for k in rt.RAY_DTYPE.names:
        if k != "tfar":
            assert np.array_equal(s2[k].view(np.uint32), fx.to_ray(g["rays"])[k].view(np.uint32)), k
    product.lib.rtcReleaseScene(sc)


def test_known_answers(product, gpu_device):
    """TriangleHitTest / SmallTriangleHitTest / WatertightTest expectations (verify.cpp:2339-3048)."""
    g = cases.load_golden("triangle_hit")
    sc, keep = build(product, gpu_device, g)
    r = g["rays"].copy()
    product.intersect(sc, r, inst_id=5)
    assert (r["geomID"] == 0).all() and (r["primID"] == 0).all() and (r["instID"] == 5).all()
    assert np.abs(r["u"] - g["expect_u"]).max() <= 16 * ULP and np.abs(r["v"] - g["expect_v"]).max() <= 16 * ULP
    assert np.abs(r["tfar"] - 1.0).max() <= 16 * ULP
    assert np.array_equal(np.stack([r["Ng_x"], r["Ng_y"], r["Ng_z"]], 1), np.tile(np.array([0, 0, 1], np.float32), (256, 1)))
    o = fx.to_ray(g["rays"])
    product.occluded(sc, o)
    assert np.isneginf(o["tfar"]).all()
    product.lib.rtcReleaseScene(sc)
    g = cases.load_golden("small_triangles")
    sc, keep = build(product, gpu_device, g)
    r = g["rays"].copy()
    product.intersect(sc, r)
    assert (r["primID"] != g["expect_prim"]).mean() <= 2e-5
    product.lib.rtcReleaseScene(sc)
    g = cases.load_golden("robust_far_sphere")
    sc, keep = build(product, gpu_device, g)
    r = g["rays"].copy()
    product.intersect(sc, r)
    assert (r["geomID"] == INV).mean() <= 2e-5
    product.lib.rtcReleaseScene(sc)


@pytest.mark.parametrize("flags", [0, rt.RTC_SCENE_FLAG_ROBUST])
def test_matches_oracle_on_seeded_scene(product, gpu_device, oracle, flags):
    meshes = fx.scene_c2(0.15)
    sc, keep = product.build_scene(gpu_device, meshes, flags)
    h = oracle.build(meshes, robust=bool(flags))
    prim = fx.primary_rays(192, 192, **fx.C2_CAMERA)
    a, b = prim.copy(), prim.copy()
    product.intersect(sc, a, coherent=True)
    oracle.intersect(h, b)
    assert parity.compare_closest(a, b)["pass"]
    for sid in (0, 1):
        d = fx.diffuse_rays(b, sample_id=sid)
        a2, b2 = d.copy(), d.copy()
        product.intersect(sc, a2)
        oracle.intersect(h, b2)
        res = parity.compare_closest(a2, b2)
        assert res["pass"], res
    s = fx.shadow_rays(b)
    s1, s2 = s.copy(), s.copy()
    product.occluded(sc, s1)
    oracle.occluded(h, s2)
    assert parity.compare_occluded(s1, s2)["pass"]
    b1, b2 = rt.Bounds(), oracle.bounds(h)
    product.lib.rtcGetSceneBounds(sc, C.byref(b1))
    assert np.array_equal(np.array([b1.lower_x, b1.lower_y, b1.lower_z, b1.upper_x, b1.upper_y, b1.upper_z], np.float32), b2)
    oracle.free(h)
    product.lib.rtcReleaseScene(sc)


def _modes(product, sc, rays):
    """IntersectWithMode (tutorials/verify/rtcore_helpers.h:751-879): one ray array through every API flavour."""
    L = product.lib
    n = len(rays)
    out = {}
    ctx = product.context()
    r = rays.copy()
    for i in range(n):
        L.rtcIntersect1(sc, C.byref(ctx), r[i:i + 1].ctypes.data)
    out["1"] = r
    r = rays.copy()
    L.rtcIntersect1M(sc, C.byref(ctx), r.ctypes.data, n, 80)
    out["1M"] = r
    r = rays.copy()
    ptrs = (C.c_void_p * n)(*[r[i:i + 1].ctypes.data for i in range(n)])
    L.rtcIntersect1Mp(sc, C.byref(ctx), ptrs, n)
    out["1Mp"] = r
    for w in (4, 8, 16):
        r = rays.copy()
        fn = getattr(L, f"rtcIntersect{w}")
        names = rt.RAYHIT_DTYPE.names
        for s in range(0, n, w):
            m = min(w, n - s)
            pk = np.zeros((20, w), dtype=np.uint32)                   # SoA packet: 12 ray fields + 8 hit fields
            for k, nm in enumerate(names):
                pk[k, :m] = r[nm][s:s + m].view(np.uint32)
            valid = np.zeros(w, dtype=np.int32)
            valid[:m] = -1
            buf = np.zeros(20 * w + 16, dtype=np.uint32)              # 64-byte aligned storage
            off = (-buf.ctypes.data % 64) // 4
            view = buf[off:off + 20 * w].reshape(20, w)
            view[:] = pk
            fn(valid.ctypes.data, sc, C.byref(ctx), view.ctypes.data)
            for k, nm in enumerate(names):
                r[nm][s:s + m] = view[k, :m].view(r[nm].dtype)
        out[str(w)] = r
    r = rays.copy()
    N = 8
    blocks = (n + N - 1) // N
    soa = np.zeros((blocks, 20, N), dtype=np.uint32)
    for k, nm in enumerate(rt.RAYHIT_DTYPE.names):
        col = np.zeros(blocks * N, dtype=np.uint32)
        col[:n] = r[nm].view(np.uint32)
        if nm == "tnear":
            col[n:] = np.float32(np.inf).view(np.uint32)               # padding lanes inactive: tnear=inf, tfar=0
        soa[:, k, :] = col.reshape(blocks, N)
    L.rtcIntersectNM(sc, C.byref(ctx), soa.ctypes.data, N, blocks, 20 * N * 4)
    for k, nm in enumerate(rt.RAYHIT_DTYPE.names):
        r[nm] = soa[:, k, :].reshape(-1)[:n].view(r[nm].dtype)
    out["NM"] = r
    return out


def test_all_api_modes_agree(product, gpu_device):
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, gpu_device, g)
    rays = g["rays"][:203].copy()
    res = _modes(product, sc, rays)
    for k, r in res.items():
        assert np.array_equal(r.view(np.uint8), res["1M"].view(np.uint8)), k
    assert parity.compare_closest(res["1M"], g["closest"][:203])["pass"]
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    product.lib.rtcReleaseScene(sc)


def test_device_resident_strided_and_unaligned_streams(product, gpu_device):
    import torch
    g = cases.load_golden("sphere_small")
    sc, keep = build(product, gpu_device, g)
    ref = g["rays"].copy()
    product.intersect(sc, ref)
    n = len(ref)
    raw = g["rays"].view(np.uint8).reshape(n, 80)
    d = torch.from_numpy(raw.copy()).cuda()                              # device-resident, stride 80
    product.intersect_ptr(sc, d.data_ptr(), n)
    assert np.array_equal(d.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE), ref)
    for stride in (96, 112, 84):                                         # BufferStrideTest; 84 = only 4-byte aligned
        host = np.full((n, stride), 0xAB, dtype=np.uint8)
        host[:, :80] = raw
        before = host.copy()
        ctx = product.context()
        product.lib.rtcIntersect1M(sc, C.byref(ctx), host.ctypes.data, n, stride)
        assert np.array_equal(host[:, :80].copy().reshape(-1).view(rt.RAYHIT_DTYPE), ref), stride
        assert np.array_equal(host[:, 80:], before[:, 80:])              # padding between records untouched
        dd = torch.from_numpy(before.copy()).cuda()
        product.lib.rtcIntersect1M(sc, C.byref(ctx), dd.data_ptr(), n, stride)
        assert np.array_equal(dd.cpu().numpy(), host), stride
    o = fx.to_ray(g["rays"])
    oref = o.copy()
    product.occluded(sc, oref)
    dd = torch.from_numpy(o.view(np.uint8).reshape(n, 48).copy()).cuda()
    product.occluded_ptr(sc, dd.data_ptr(), n)
    assert np.array_equal(dd.cpu().numpy().reshape(-1).view(rt.RAY_DTYPE), oref)
    product.lib.rtcReleaseScene(sc)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pinned_host_streams(product, mode):
    """Page-locked host streams: staged both ways (zerocopy=0, default), staged in + hit fields written by the
    kernel over PCIe (1), traced in place over PCIe (2).  All must be bit-identical to the pageable
    path, with the padding between records untouched."""
    import torch
    dev = product.new_device(f"zerocopy={mode}")
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, dev, g)
    reps = 4096 // len(g["rays"]) + 2                                    # the zero-copy paths start at 4096 rays
    rays = np.tile(g["rays"], reps)
    n = len(rays)
    ref = rays.copy()
    product.intersect(sc, ref)                                           # pageable -> staged copies
    for stride in (80, 96):
        host = np.full((n, stride), 0xCD, dtype=np.uint8)
        host[:, :80] = rays.view(np.uint8).reshape(n, 80)
        pinned = torch.from_numpy(host.copy()).pin_memory()
        ctx = product.context()
        product.lib.rtcIntersect1M(sc, C.byref(ctx), pinned.data_ptr(), n, stride)
        out = pinned.numpy()
        assert np.array_equal(out[:, :80].copy().reshape(-1).view(rt.RAYHIT_DTYPE), ref), stride
        assert np.array_equal(out[:, 80:], host[:, 80:])
    o = fx.to_ray(rays)
    oref = o.copy()
    product.occluded(sc, oref)
    pinned = torch.from_numpy(o.view(np.uint8).reshape(n, 48).copy()).pin_memory()
    product.occluded_ptr(sc, pinned.data_ptr(), n)
    assert np.array_equal(pinned.numpy().reshape(-1).view(rt.RAY_DTYPE), oref)
    assert product.lib.rtcGetDeviceError(dev) == 0
    product.lib.rtcReleaseScene(sc)
    product.lib.rtcReleaseDevice(dev)


def test_empty_scenes_updates_and_disable(product, gpu_device):
    """EmptySceneTest :1061, EmptyGeometryTest :1093, UpdateTest :1710, enable/disable, detach."""
    L = product.lib
    ctx = product.context()
    sc = L.rtcNewScene(gpu_device)
    L.rtcCommitScene(sc)
    r = fx.incoherent_rays(100)
    before = r.copy()
    product.intersect(sc, r)
    assert np.array_equal(r, before)
    g0 = L.rtcNewGeometry(gpu_device, rt.RTC_GEOMETRY_TYPE_TRIANGLE)       # geometry without buffers
    L.rtcCommitGeometry(g0)
    L.rtcAttachGeometry(sc, g0)
    L.rtcCommitScene(sc)
    product.intersect(sc, r)
    assert np.array_equal(r, before) and L.rtcGetDeviceError(gpu_device) == 0
    keep = []
    v, t = fx.triangle_sphere((0, 0, 0), 1.0, 8)
    gid, g1 = product.add_mesh(gpu_device, sc, v, t, keep)
    assert gid == 1
    product.intersect(sc, r)                                               # attached but not committed
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    L.rtcCommitScene(sc)
    r1 = before.copy()
    product.intersect(sc, r1)
    assert (r1["geomID"] == 1).all()
    keep[0][:v.size] *= 2.0                                                # UpdateTest: scale the shared vertex buffer
    L.rtcUpdateGeometryBuffer(g1, rt.RTC_BUFFER_TYPE_VERTEX, 0)
    L.rtcCommitGeometry(g1)
    L.rtcCommitScene(sc)
    r2 = before.copy()
    product.intersect(sc, r2)
    assert np.allclose(r2["tfar"], 2 * r1["tfar"], rtol=1e-5)
    L.rtcDisableGeometry(g1)
    L.rtcCommitScene(sc)
    r3 = before.copy()
    product.intersect(sc, r3)
    assert np.array_equal(r3, before)
    L.rtcEnableGeometry(g1)
    L.rtcCommitScene(sc)
    r4 = before.copy()
    product.intersect(sc, r4)
    assert np.array_equal(r4, r2)
    L.rtcDetachGeometry(sc, 1)
    L.rtcCommitScene(sc)
    r5 = before.copy()
    product.intersect(sc, r5)
    assert np.array_equal(r5, before) and L.rtcGetDeviceError(gpu_device) == 0
    L.rtcReleaseGeometry(g0)
    L.rtcReleaseGeometry(g1)
    L.rtcReleaseScene(sc)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_hit_download_modes(product, mode):
    """Host-staged streams with the alternative hit downloads (d2h=1/2: strided 2-D copy of the hit part, d2h=3: compact
    hit list scattered by a host thread) must equal the default path bit for bit, pageable and page-locked, for several
    chunkings, strides and both query kinds; everything outside tfar / hit stays untouched."""
    import torch
    g = cases.load_golden("two_geoms")
    base = product.new_device("d2h=0")                                   # whole-span download: the plain path
    sc0, keep0 = build(product, base, g)
    reps = 200000 // len(g["rays"]) + 1                                  # > 65536 rays: the compact path engages
    rays = np.tile(g["rays"], reps)
    rays["id"] = np.arange(len(rays), dtype=np.uint32)
    n = len(rays)
    ref = rays.copy()
    product.intersect(sc0, ref)
    oref = fx.to_ray(rays)
    product.occluded(sc0, oref)
    cfgs = [f"d2h={mode}", f"d2h={mode},chunk_rays=50000", f"d2h={mode},chunk_rays=70001"]
    if mode == 3:                                                        # compact_min_rays: engage the compact path on this 200 K-ray stream
        cfgs = [c + ",compact_min_rays=0" for c in cfgs] + ["d2h=3"]
        cfgs += [c + ",compact_min_rays=0" for c in ("d2h=3,pack_rays=1", "d2h=3,pack_rays=2,host_threads=3,chunk_rays=33333", "d2h=3,host_threads=1")]
    for cfg in cfgs:
        dev = product.new_device(cfg)
        sc, keep = build(product, dev, g)
        a = rays.copy()
        product.intersect(sc, a)                                         # pageable host memory
        assert np.array_equal(a, ref), cfg
        for stride in (80, 112):
            host = np.full((n, stride), 0xAB, dtype=np.uint8)
            host[:, :80] = rays.view(np.uint8).reshape(n, 80)
            pinned = torch.from_numpy(host.copy()).pin_memory()
            ctx = product.context()
            product.lib.rtcIntersect1M(sc, C.byref(ctx), pinned.data_ptr(), n, stride)
            out = pinned.numpy()
            assert np.array_equal(out[:, :80].copy().reshape(-1).view(rt.RAYHIT_DTYPE), ref), (cfg, stride)
            assert np.array_equal(out[:, 80:], host[:, 80:])
        o = fx.to_ray(rays)
        product.occluded(sc, o)
        assert np.array_equal(o, oref), cfg
        assert product.lib.rtcGetDeviceError(dev) == 0
        product.lib.rtcReleaseScene(sc)
        product.lib.rtcReleaseDevice(dev)
    product.lib.rtcReleaseScene(sc0)
    product.lib.rtcReleaseDevice(base)


def _wave(v0, phase):
    """Deformation used by the refit tests: a travelling wave on y plus a drift on x (float32 throughout)."""
    v = v0.copy()
    v[:, 1] += (0.35 * np.sin(1.3 * v0[:, 0] + phase) * np.cos(0.9 * v0[:, 2] - 0.5 * phase)).astype(np.float32)
    v[:, 0] += np.float32(0.05 * phase)
    return v


def test_refit_matches_fresh_build(product, gpu_device, oracle):
    """RTC_BUILD_QUALITY_REFIT geometries whose vertex buffer is updated are refitted
    (topology kept, boxes re-quantised bottom-up); answers must equal those of a BVH built from
    scratch over the new positions (oracle), including dropped (NaN) triangles and scene bounds."""
    L = product.lib
    base = [fx.displaced_plane(96, extent=4.0), fx.triangle_sphere((0.3, 1.2, -0.2), 0.8, 40)]
    sc = L.rtcNewScene(gpu_device)
    keep, geoms = [], []
    for v, t in base:
        gid, g = product.add_mesh(gpu_device, sc, v, t, keep)
        L.rtcSetGeometryBuildQuality(g, rt.RTC_BUILD_QUALITY_REFIT)
        L.rtcCommitGeometry(g)
        geoms.append(g)
    L.rtcCommitScene(sc)
    st0 = product.build_stats(sc)
    assert st0["refitCount"] == 0
    rays = np.concatenate([fx.incoherent_rays(40000, org=(0.1, 2.5, 0.2), seed=3),
                           fx.primary_rays(160, 160, org=(0.5, 6.0, 0.5), look=(0, -1, 0), up=(0, 0, 1))])
    vbufs = [keep[0], keep[2]]                                             # the shared (padded) vertex arrays
    for step, phase in enumerate((0.7, 1.9, 3.1), start=1):
        cur = []
        for (v0, t), buf in zip(base, vbufs):
            v = _wave(np.asarray(v0, dtype=np.float32), phase)
            if step == 2 and len(v) > 1000:
                v[123] = np.nan                                            # triangles using this vertex vanish (scene_triangle_mesh.h:131-153)
            buf[:v.size] = v.ravel()
            cur.append((v, t))
        for g in geoms:
            L.rtcUpdateGeometryBuffer(g, rt.RTC_BUFFER_TYPE_VERTEX, 0)
            L.rtcCommitGeometry(g)
        L.rtcCommitScene(sc)
        st = product.build_stats(sc)
        assert st["refitCount"] == step and st["numNodes"] == st0["numNodes"] and st["numTris"] == st0["numTris"], st
        assert 1.0 < st["sahExact"] <= st["sah"] < 200.0
        h = oracle.build(cur)
        b = rt.Bounds()
        L.rtcGetSceneBounds(sc, C.byref(b))
        ours_b = np.array([b.lower_x, b.lower_y, b.lower_z, b.upper_x, b.upper_y, b.upper_z], dtype=np.float32)
        assert np.array_equal(ours_b, oracle.bounds(h)), (ours_b, oracle.bounds(h))
        a, w = rays.copy(), rays.copy()
        product.intersect(sc, a)
        oracle.intersect(h, w)
        res = parity.compare_closest(a, w)
        assert res["pass"] and res["hits_ours"] > 10000, res
        sa = fx.shadow_rays(w)
        sw = sa.copy()
        product.occluded(sc, sa)
        oracle.occluded(h, sw)
        assert parity.compare_occluded(sa, sw)["pass"]
        oracle.free(h)
    # a topology change falls back to a full build
    L.rtcDisableGeometry(geoms[1])
    L.rtcCommitScene(sc)
    assert product.build_stats(sc)["refitCount"] == 0
    # MEDIUM quality geometry is rebuilt, never refitted
    L.rtcEnableGeometry(geoms[1])
    L.rtcSetGeometryBuildQuality(geoms[0], rt.RTC_BUILD_QUALITY_MEDIUM)
    L.rtcCommitGeometry(geoms[0])
    L.rtcCommitScene(sc)
    L.rtcUpdateGeometryBuffer(geoms[0], rt.RTC_BUFFER_TYPE_VERTEX, 0)
    L.rtcCommitGeometry(geoms[0])
    L.rtcCommitScene(sc)
    assert product.build_stats(sc)["refitCount"] == 0 and L.rtcGetDeviceError(gpu_device) == 0
    for g in geoms:
        L.rtcReleaseGeometry(g)
    L.rtcReleaseScene(sc)


def test_scene_build_quality_low_and_high(product, gpu_device, oracle):
    """RTC_BUILD_QUALITY_LOW / HIGH select the fast / the thorough front end; answers do not change."""
    L = product.lib
    meshes = [fx.displaced_plane(64, extent=3.0), fx.triangle_sphere((0, 1, 0), 0.7, 24)]
    h = oracle.build(meshes)
    rays = fx.incoherent_rays(30000, org=(0.2, 2.0, 0.1), seed=11)
    want = rays.copy()
    oracle.intersect(h, want)
    sah = {}
    for q in (rt.RTC_BUILD_QUALITY_LOW, rt.RTC_BUILD_QUALITY_MEDIUM, rt.RTC_BUILD_QUALITY_HIGH):
        keep = []
        sc = L.rtcNewScene(gpu_device)
        L.rtcSetSceneBuildQuality(sc, q)
        for v, t in meshes:
            _, g = product.add_mesh(gpu_device, sc, v, t, keep)
            L.rtcReleaseGeometry(g)
        L.rtcCommitScene(sc)
        st = product.build_stats(sc)
        sah[q] = st["sah"]
        assert (st["builderIterations"] == 0) == (q == rt.RTC_BUILD_QUALITY_LOW) or q == rt.RTC_BUILD_QUALITY_MEDIUM
        a = rays.copy()
        product.intersect(sc, a)
        assert parity.compare_closest(a, want)["pass"]
        L.rtcReleaseScene(sc)
    assert sah[rt.RTC_BUILD_QUALITY_HIGH] <= sah[rt.RTC_BUILD_QUALITY_LOW] * 1.02
    oracle.free(h)


def test_concurrent_queries_from_threads(product, gpu_device):
    """Queries are re-entrant on a committed scene (SURVEY 8b threading)."""
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, gpu_device, g)
    outs = [g["rays"].copy() for _ in range(4)]
    th = [threading.Thread(target=lambda r=r: [product.intersect(sc, r) for _ in range(1)]) for r in outs]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in outs:
        assert np.array_equal(r, outs[0])
    assert parity.compare_closest(outs[0], g["closest"])["pass"]
    product.lib.rtcReleaseScene(sc)


def test_many_threads_small_tiles(product, gpu_device):
    """Tile-sized calls from many threads (how the reference's stream tutorials drive rtcIntersect1M): every call takes its own
    staging context, results equal the single big call, no errors."""
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, gpu_device, g)
    rays = np.tile(g["rays"], 12)                                          # 72 000 rays
    want = rays.copy()
    product.intersect(sc, want)
    got = rays.copy()
    occ_want = fx.to_ray(rays)
    product.occluded(sc, occ_want)
    occ_got = fx.to_ray(rays)
    tile = 1024
    tiles = [(i, min(i + tile, len(rays))) for i in range(0, len(rays), tile)]

    def worker(k, nthreads):
        for b, e in tiles[k::nthreads]:
            product.intersect(sc, got[b:e])
            product.occluded(sc, occ_got[b:e])
    nthreads = 12
    th = [threading.Thread(target=worker, args=(k, nthreads)) for k in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert np.array_equal(got, want) and np.array_equal(occ_got, occ_want)
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    product.lib.rtcReleaseScene(sc)


def test_image_roundtrip_replica(product, gpu_device):
    """A byte copy of the flat BVH image is a usable replica (what the NVLink broadcast ships)."""
    import torch
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, gpu_device, g)
    nbytes = C.c_size_t()
    ptr = product.lib.rtcxGetSceneImage(sc, C.byref(nbytes))
    assert ptr and nbytes.value % 128 == 0
    sc2 = product.lib.rtcNewScene(gpu_device)
    product.lib.rtcxSetSceneImage(sc2, ptr, nbytes.value)
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    a, b = g["rays"].copy(), g["rays"].copy()
    product.intersect(sc, a)
    product.lib.rtcReleaseScene(sc)                                        # the replica owns its own copy
    product.intersect(sc2, b)
    assert np.array_equal(a, b)
    st = product.build_stats(sc2)
    assert st["numTris"] == fx.num_tris(g["meshes"])
    bad = torch.zeros(256, dtype=torch.uint8, device="cuda")
    product.lib.rtcxSetSceneImage(sc2, bad.data_ptr(), 256)
    assert product.lib.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_ARGUMENT
    product.lib.rtcReleaseScene(sc2)


def test_image_save_and_load(product, gpu_device, tmp_path):
    """the flat image written to a file is a loadable BVH."""
    g = cases.load_golden("two_geoms")
    sc, keep = build(product, gpu_device, g)
    path = str(tmp_path / "scene.rqb").encode()
    assert product.lib.rtcxSaveSceneImage(sc, path) == 0
    sc2 = product.lib.rtcNewScene(gpu_device)
    assert product.lib.rtcxLoadSceneImage(sc2, path) == 0
    a, b = g["rays"].copy(), g["rays"].copy()
    product.intersect(sc, a)
    product.intersect(sc2, b)
    assert np.array_equal(a, b) and parity.compare_closest(b, g["closest"])["pass"]
    with open(path.decode(), "r+b") as f:                                  # corrupt the magic: rejected, scene keeps its image
        f.write(b"garbage!")
    assert product.lib.rtcxLoadSceneImage(sc2, path) == -1
    assert product.lib.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_ARGUMENT
    assert product.lib.rtcxLoadSceneImage(sc2, b"/nonexistent/dir/x.rqb") == -1
    assert product.lib.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_ARGUMENT
    c = g["rays"].copy()
    product.intersect(sc2, c)
    assert np.array_equal(a, c)
    product.lib.rtcReleaseScene(sc)
    product.lib.rtcReleaseScene(sc2)


def test_build_statistics_and_counters(product, gpu_device):
    meshes = fx.scene_c1()
    sc, keep = product.build_scene(gpu_device, meshes)
    st = product.build_stats(sc)
    assert st["numPrimsIn"] == 32760 and st["numPrimsValid"] == 32760 and st["numTris"] == 32760
    assert 1 <= st["depth"] <= 32 and st["numNodes"] < 32760 / 2
    assert 5.0 < st["sahExact"] <= st["sah"] < 30.0          # reference's own tree: 12.16
    r = fx.incoherent_rays(1 << 16, seed=9)
    c = product.intersect_counted(sc, r)
    assert c["rays"] == len(r) and c["hits"] == len(r)
    assert 3 < c["nodes"] / c["rays"] < 40 and 1 <= c["tris"] / c["rays"] < 20 and c["stackMax"] <= st["depth"]
    r2 = fx.incoherent_rays(1 << 16, seed=9)
    product.intersect(sc, r2)
    assert np.array_equal(r, r2)                             # instrumented and fast kernels agree bit for bit
    product.lib.rtcReleaseScene(sc)


def test_staged_geometry_upload_equals_plain_copy(product):
    """Large pageable vertex / index buffers reach the GPU through the page-locked chunk ring (rtcore_api.cpp::stagedUpload; here
    from 4 MB on instead of the default 32): same answers as the plain cudaMemcpyAsync route, for a buffer shorter than one chunk
    (6.6 MB) and one of 2.4 chunks (19.9 MB)."""
    res = []
    for cfg in ("stage_geometry=0", "stage_geometry=4"):
        dev = product.new_device(cfg)
        out = []
        for scale in (0.75, 1.3):                            # 0.56 M / 1.69 M triangles
            sc, keep = product.build_scene(dev, fx.scene_c2(scale))
            r = fx.incoherent_rays(1 << 17, org=(0.3, 6.0, -0.2), seed=11)
            product.intersect(sc, r)
            out.append(r)
            product.lib.rtcReleaseScene(sc)
        assert product.lib.rtcGetDeviceError(dev) == 0
        res.append(out)
        product.lib.rtcReleaseDevice(dev)
    for a, b in zip(res[0], res[1]):
        assert (a["geomID"] != 0xFFFFFFFF).sum() > 1000
        assert np.array_equal(a, b)


def test_full_size_c2_properties(product, gpu_device, oracle):
    """BASELINE config 1 at full size: ~1.0 M triangles, 4096x4096 primary -> 16.7 M diffuse + shadow rays.
    Checked through size-independent properties plus an oracle comparison on a seeded subsample."""
    import torch
    meshes = fx.scene_c2(1.0)
    assert 990000 < fx.num_tris(meshes) < 1010000
    sc, keep = product.build_scene(gpu_device, meshes)
    st = product.build_stats(sc)
    assert st["numTris"] == fx.num_tris(meshes)
    hits_total, rays_total = 0, 0
    sub_in, sub_out = [], []
    for band in range(8):                                                  # 8 bands of 512 rows keep host memory modest
        prim = fx.primary_rays(4096, 4096, rows=(band * 512, band * 512 + 512), **fx.C2_CAMERA)
        product.intersect(sc, prim, coherent=True)
        assert (prim["geomID"] != INV).mean() > 0.9999                     # the camera sees only geometry (rare edge leaks, as in the reference)
        d = fx.diffuse_rays(prim)
        d_in = d.copy()
        product.intersect(sc, d)
        hit = d["geomID"] != INV
        rays_total += len(d)
        hits_total += int(hit.sum())
        # property 1: hit records are consistent -- barycentrics inside, t inside the segment, ids in range
        assert (d["u"][hit] >= 0).all() and (d["v"][hit] >= 0).all() and (d["u"][hit] + d["v"][hit] <= 1 + 1e-6).all()
        assert (d["tfar"][hit] > d["tnear"][hit]).all() and (d["geomID"][hit] <= 1).all()
        assert np.array_equal(d[~hit], d_in[~hit])                         # misses untouched
        # property 2 ("intersect then occluded agree", rtcore_helpers.h:881-904): a segment reaching just past
        # the hit is occluded, a segment stopping well short of it is not; missing rays are never occluded
        o = fx.to_ray(d_in)
        o["tfar"][hit] = d["tfar"][hit] * np.float32(1.001)
        product.occluded(sc, o)
        assert np.isneginf(o["tfar"][hit]).all() and not np.isneginf(o["tfar"][~hit]).any()
        o = fx.to_ray(d_in)
        o["tfar"][hit] = d["tfar"][hit] * np.float32(0.999)
        o["tfar"][~hit] = 50.0
        product.occluded(sc, o)
        assert np.isneginf(o["tfar"]).sum() == 0
        # property 3: idempotence -- retracing a segment that ends just past the found distance returns the same hit
        again = d.copy()
        again["tfar"][hit] *= np.float32(1.000001)
        again["geomID"] = INV
        product.intersect(sc, again)
        same = (again["primID"] == d["primID"]) & (again["geomID"] == d["geomID"])
        assert same[hit].mean() > 0.9999
        sel = np.arange(band * 7, len(d_in), 997)[:800]
        sub_in.append(d_in[sel])
        sub_out.append(d[sel])
    assert 0.9999 * 4096 * 4096 <= rays_total <= 4096 * 4096 and 0.1 < hits_total / rays_total < 0.9
    # oracle on the subsample (6 400 rays against the full 1 M-triangle scene)
    h = oracle.build(meshes)
    a = np.concatenate(sub_out)
    b = np.concatenate(sub_in)
    oracle.intersect(h, b)
    res = parity.compare_closest(a, b)
    oracle.free(h)
    assert res["pass"], res
    product.lib.rtcReleaseScene(sc)
