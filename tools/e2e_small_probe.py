"""Short-stream end-to-end probe: configs[0] (1 M coherent primary rays on the 32 760-triangle sphere) through rtcIntersect1M on a
page-locked host buffer, per device configuration.  usage: python tools/e2e_small_probe.py "<cfg>" ..."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures
lib = rt.RTCore()
meshes = fx.scene_c1()
rays = fx.primary_rays(1024, 1024, org=(0.0, 0.0, -3.0), look=(0, 0, 1), up=(0, 1, 0))
n = len(rays)
host = torch.from_numpy(rays.view(np.uint8).reshape(n, 80).copy()).pin_memory()
hw = torch.empty_like(host).pin_memory()
first = None
for cfg in sys.argv[1:] or ["-"]:
    dev = lib.new_device("" if cfg == "-" else cfg)
    sc, keep = lib.build_scene(dev, meshes)
    ts = []
    for r in range(8):
        hw.copy_(host); torch.cuda.synchronize()
        time.sleep(float(os.environ.get("PROBE_SETTLE", "0")))          # optional settle time after the host-side reset copy (PROBE_SETTLE seconds): measured, no consistent effect
        t0 = time.perf_counter(); lib.intersect_ptr(sc, hw.data_ptr(), n, coherent=True); ts.append(time.perf_counter() - t0)
    if first is None:
        first = hw.numpy().copy()
    print(f"{cfg:40s} best {min(ts[2:]) * 1e3:6.2f} ms  median {np.median(ts[2:]) * 1e3:6.2f} ms  {n / min(ts[2:]) / 1e6:7.1f} Mrays/s  same={bool(np.array_equal(first, hw.numpy()))}", flush=True)
    lib.lib.rtcReleaseScene(sc); lib.lib.rtcReleaseDevice(dev)
