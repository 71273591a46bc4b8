"""First GPU run: build + trace the small configurations with the product and the reference, compare."""
import importlib, json, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore
parity = importlib.import_module("embree-aarch64_b200.parity")
import torch

ours = rt.RTCore()
ref = rt.RTCore("oracle/_ref/libembree3_ref.so")
dO = ours.new_device("verbose=1,benchmark=1")
dR = ref.new_device("")

def run(name, meshes, rays, coherent=False, flags=0):
    t0 = time.time(); sO, kO = ours.build_scene(dO, meshes, flags); tO = time.time() - t0
    t0 = time.time(); sR, kR = ref.build_scene(dR, meshes, flags); tR = time.time() - t0
    print(name, "tris", fx.num_tris(meshes), "commit ours %.3fs ref %.3fs" % (tO, tR), ours.build_stats(sO), flush=True)
    a, b = rays.copy(), rays.copy()
    t0 = time.time(); ours.intersect(sO, a, coherent); t1 = time.time(); ref.intersect(sR, b, coherent); t2 = time.time()
    print("  closest: ours(host path) %.3fs ref(1 thread) %.3fs" % (t1 - t0, t2 - t1), json.dumps(parity.compare_closest(a, b)), flush=True)
    c = ours.intersect_counted(sO, rays.copy())
    print("  counters", c, "nodes/ray %.2f tris/ray %.2f" % (c["nodes"] / max(c["rays"], 1), c["tris"] / max(c["rays"], 1)), flush=True)
    # device-resident timing
    d = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), -1).copy()).cuda()
    for _ in range(2):
        d2 = d.clone(); torch.cuda.synchronize(); t0 = time.time()
        ours.intersect_ptr(sO, d2.data_ptr(), len(rays)); torch.cuda.synchronize(); dt = time.time() - t0
    back = d2.cpu().numpy().view(rt.RAYHIT_DTYPE).reshape(-1)
    print("  device-resident: %.3f ms  %.1f Mrays/s  identical to host path: %s" % (dt * 1e3, len(rays) / dt / 1e6, np.array_equal(back, a)), flush=True)
    sh = fx.shadow_rays(a)
    s1, s2 = sh.copy(), sh.copy()
    ours.occluded(sO, s1); ref.occluded(sR, s2)
    print("  occluded:", json.dumps(parity.compare_occluded(s1, s2)), flush=True)
    ours.lib.rtcReleaseScene(sO); ref.lib.rtcReleaseScene(sR)
    return a

m1 = fx.scene_c1()
a = run("C1 coherent", m1, fx.primary_rays(1024, 1024, **fx.C1_CAMERA), True)
run("C1 incoherent", m1, fx.incoherent_rays(1 << 20, seed=1))
run("C1 robust", m1, fx.incoherent_rays(1 << 18, seed=2), flags=rt.RTC_SCENE_FLAG_ROBUST)
m2 = fx.scene_c2(1.0)
prim = run("C2 primary", m2, fx.primary_rays(2048, 2048, **fx.C2_CAMERA), True)
run("C2 diffuse", m2, fx.diffuse_rays(prim))
print("launches", ours.lib.rtcxGetLaunchCount())
