"""Every flavour of the query API: rtcIntersect1 / 1M / 1Mp / 4 / 8 / 16 / NM / Np and the rtcOccluded twins,
layouts in host memory and -- for the product -- in GPU memory (gather / trace / scatter stay on the device), all funnelled
through one harness (tests/modes.py = the reference's IntersectWithMode, tutorials/verify/rtcore_helpers.h:751-904).
The harness itself is proven against the real reference library where oracle/_ref exists (CPU test)."""
import ctypes as C

import numpy as np
import pytest

import cases
import modes

parity = cases.importlib.import_module("embree-aarch64_b200.parity")
rt, fx = cases.rt, cases.fx
INV = 0xFFFFFFFF


def test_mode_harness_against_the_reference_library(reflib):
    """Proves the harness: the reference library answers every entry point identically to its golden rtcIntersect1M /
    rtcOccluded1M vectors (packets may pick another of two coincident-t triangles: ids compared through parity rules)."""
    g = cases.load_golden("two_geoms")
    dev = reflib.new_device("threads=2")
    sc, keep = reflib.build_scene(dev, g["meshes"], g["flags"])
    rays = g["rays"][:203].copy()
    for m in modes.MODES:
        a = modes.run_mode(reflib, sc, rays, m, occluded=False)
        assert parity.compare_closest(a, g["closest"][:203])["pass"], m
        b = modes.run_mode(reflib, sc, fx.to_ray(rays), m, occluded=True)
        assert parity.compare_occluded(b, g["occl_self_out"][:203])["disagree"] == 0, m
    assert reflib.lib.rtcGetDeviceError(dev) == 0
    reflib.lib.rtcReleaseScene(sc)
    reflib.lib.rtcReleaseDevice(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
def test_every_entry_point_closest_and_occluded(product, gpu_device, device):
    g = cases.load_golden("two_geoms")
    sc, keep = product.build_scene(gpu_device, g["meshes"], g["flags"])
    rays = g["rays"][:203].copy()
    ref_c = modes.run_mode(product, sc, rays, "1M", False)
    ref_o = modes.run_mode(product, sc, fx.to_ray(rays), "1M", True)
    assert parity.compare_closest(ref_c, g["closest"][:203])["pass"]
    assert parity.compare_occluded(ref_o, g["occl_self_out"][:203])["disagree"] == 0
    launches0 = product.lib.rtcxGetLaunchCount()
    for m in modes.MODES:
        if device and m == "1":
            continue                                                     # a single device-resident ray is the 1M case with M = 1
        a, b, agree = modes.intersect_then_occluded_agree(product, sc, rays, m, device)
        assert np.array_equal(a.view(np.uint8), ref_c.view(np.uint8)), (m, device)
        assert np.array_equal(b.view(np.uint8), ref_o.view(np.uint8)), (m, device)
        assert agree, (m, device)                                        # closest hit found <=> occluded, per ray
    assert product.lib.rtcGetDeviceError(gpu_device) == 0
    assert product.lib.rtcxGetLaunchCount() > launches0
    product.lib.rtcReleaseScene(sc)


@pytest.mark.gpu
def test_nm_and_np_are_one_launch_per_call(product, gpu_device):
    """rtcIntersectNM / Np over all N x M rays = one gather + one trace + one scatter (device layouts), or one staged trace
    (host layouts) -- never one launch per packet (round-1 behaviour)."""
    g = cases.load_golden("sphere_small")
    sc, keep = product.build_scene(gpu_device, g["meshes"], g["flags"])
    rays = g["rays"][:4096].copy()
    ref = modes.run_mode(product, sc, rays, "1M", False)
    for device in (False, True):
        for m in ("NM", "Np"):
            l0 = product.lib.rtcxGetLaunchCount()
            a = modes.run_mode(product, sc, rays, m, False, device)
            dl = product.lib.rtcxGetLaunchCount() - l0
            assert np.array_equal(a.view(np.uint8), ref.view(np.uint8)), (m, device)
            assert dl <= (3 if device else 1), (m, device, dl)
    product.lib.rtcReleaseScene(sc)


@pytest.mark.gpu
def test_packet_lanes_follow_packet_entry_rules(product, gpu_device, reflib):
    """Occlusion rays with tnear < 0: the stream filters (1M, NM with N = the SIMD width, Np: octant-sorting branches of
    stream_filters.cpp -> occludedN) skip them (bvh_intersector_stream.cpp:303-305) while single rays and the packet kernels
    clamp tnear to 0 and test the ray (bvh_intersector1.cpp:132, bvh_intersector_hybrid.cpp:153,403).  Compared live with
    the reference (built for AVX2: packets of 8 are its native SoA width)."""
    g = cases.load_golden("sphere_small")
    rays = fx.to_ray(g["rays"][:64].copy())
    rays["tnear"] = -1.0
    rdev = reflib.new_device("threads=1")
    rsc, rkeep = reflib.build_scene(rdev, g["meshes"], g["flags"])
    sc, keep = product.build_scene(gpu_device, g["meshes"], g["flags"])
    for m in ("1", "1M", "4", "8", "16", "NM", "Np"):
        want = modes.run_mode(reflib, rsc, rays, m, occluded=True)
        got = modes.run_mode(product, sc, rays, m, occluded=True)
        assert np.array_equal(np.isneginf(got["tfar"]), np.isneginf(want["tfar"])), m
    want4 = modes.run_mode(reflib, rsc, rays, "4", occluded=True)
    assert np.isneginf(want4["tfar"]).any()                                # the case is not vacuous: packets do test these rays
    reflib.lib.rtcReleaseScene(rsc)
    reflib.lib.rtcReleaseDevice(rdev)
    product.lib.rtcReleaseScene(sc)


@pytest.mark.gpu
def test_misaligned_and_overlapping_streams_are_rejected(product, gpu_device):
    """A stride or pointer that is not a multiple of 4 would fault inside the kernel (and poison the CUDA context); the
    boundary rejects it like the reference's debug checks do (rtcore.cpp:602-606)."""
    g = cases.load_golden("sphere_small")
    sc, keep = product.build_scene(gpu_device, g["meshes"], g["flags"])
    L = product.lib
    ctx = product.context()
    buf = np.zeros(82 * 16 + 8, dtype=np.uint8)
    L.rtcIntersect1M(sc, C.byref(ctx), buf.ctypes.data, 16, 82)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_ARGUMENT
    L.rtcIntersect1M(sc, C.byref(ctx), buf.ctypes.data + 2, 8, 80)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_ARGUMENT
    L.rtcIntersect1M(sc, C.byref(ctx), buf.ctypes.data, 8, 40)
    assert L.rtcGetDeviceError(gpu_device) == rt.RTC_ERROR_INVALID_OPERATION
    r = g["rays"][:32].copy()
    product.intersect(sc, r)                                               # the device still works
    assert L.rtcGetDeviceError(gpu_device) == 0 and (r["geomID"] != INV).any()
    L.rtcReleaseScene(sc)
