"""Answers of an experiment build of the library against the default build on the same scene and streams (bit for bit).
usage: python tools/check_variant.py <variant .so> [device cfg]"""
import importlib, sys
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore
import torch
var, cfg = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
res = []
for path in (None, var):
    lib = rt.RTCore(path) if path else rt.RTCore()
    dev = lib.new_device(cfg)
    meshes = fx.scene_c2(0.5)
    sc, keep = lib.build_scene(dev, meshes)
    prim = fx.primary_rays(1024, 1024, **fx.C2_CAMERA)
    lib.intersect(sc, prim, coherent=True)
    d = fx.diffuse_rays(prim)
    t = torch.from_numpy(d.view(np.uint8).reshape(len(d), 80).copy()).cuda()
    lib.intersect_ptr(sc, t.data_ptr(), len(d))
    torch.cuda.synchronize()
    dev_out = t.cpu().numpy().copy()
    h = np.tile(d, 5)                                       # 5 M rays from host memory: the staged LIST kernels
    lib.intersect(sc, h)
    res.append((prim.copy(), dev_out, h[:len(d)].copy(), lib.lib.rtcGetDeviceError(dev)))
same_p = np.array_equal(res[0][0].view(np.uint8), res[1][0].view(np.uint8))
same_d = np.array_equal(res[0][1], res[1][1])
same_h = np.array_equal(res[0][2].view(np.uint8), res[1][2].view(np.uint8))
host_eq_dev = np.array_equal(res[1][2].view(np.uint8).reshape(-1, 80), res[1][1])
print("variant == default: primary", same_p, "diffuse(device)", same_d, "diffuse(host-staged)", same_h, "| variant host == device", host_eq_dev, "| errors", res[0][3], res[1][3],
      "| hits", int((res[1][2]["geomID"] != 0xFFFFFFFF).sum()))
