"""BASELINE.json configs[0]: tessellated sphere (32 760 triangles), 1024x1024 coherent primary rays from (0,0,-3),
rtcIntersect1M with RTC_INTERSECT_CONTEXT_FLAG_COHERENT -- ours (device-resident stream, CUDA events; host stream, wall clock)
and the reference library on the host cores, with the parity of the two answers.
usage: python tools/bench_c1.py [--reps 10] [--no-reference]"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures
parity = importlib.import_module("embree-aarch64_b200.parity")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    import torch
    meshes = fx.scene_c1()
    rays = fx.primary_rays(1024, 1024, org=(0.0, 0.0, -3.0), look=(0, 0, 1), up=(0, 1, 0))
    n = len(rays)
    lib = rt.RTCore()
    dev = lib.new_device("async=1")
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    lib.lib.rtcxSetDeviceStream(dev, st.cuda_stream)
    sc, keep = lib.build_scene(dev, meshes)
    bs = lib.build_stats(sc)
    pristine = torch.from_numpy(rays.view(np.uint8).reshape(n, 80).copy()).cuda()
    work = pristine.clone()
    res = {"workload": f"configs[0]: {fx.num_tris(meshes)}-triangle sphere, {n} coherent primary rays", "build_ms": bs["msTotal"]}
    for name, coh in (("coherent_flag", True), ("incoherent_flag", False)):
        ts = []
        for r in range(a.reps + 3):
            work.copy_(pristine)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lib.intersect_ptr(sc, work.data_ptr(), n, coherent=coh); e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ts.append(e0.elapsed_time(e1))
        res[f"ours_device_mrays_per_s_{name}"] = n / float(np.median(ts)) / 1e3
    ours = work.cpu().numpy().reshape(-1).view(rt.RAYHIT_DTYPE)
    host = torch.from_numpy(rays.view(np.uint8).reshape(n, 80).copy()).pin_memory()
    hw = torch.empty_like(host).pin_memory()
    best = 1e9
    for r in range(5):
        hw.copy_(host); torch.cuda.synchronize()
        t0 = time.perf_counter(); lib.intersect_ptr(sc, hw.data_ptr(), n, coherent=True); best = min(best, time.perf_counter() - t0)
    res["ours_e2e_mrays_per_s"] = n / best / 1e6
    res["hit_fraction"] = float((ours["geomID"] != 0xFFFFFFFF).mean())
    if not a.no_reference:
        from oracle.rq_oracle import REF_LIB
        if os.path.exists(REF_LIB):
            ref = rt.RTCore(REF_LIB)
            nth = os.cpu_count() or 1
            rdev = ref.new_device(f"threads={nth}")
            rsc, rkeep = ref.build_scene(rdev, meshes)
            rr = rays.copy()
            rows = np.array_split(np.arange(n), nth * 8)

            def run_all():
                nxt = [0]; lock = threading.Lock()

                def w():
                    while True:
                        with lock:
                            i = nxt[0]; nxt[0] += 1
                        if i >= len(rows):
                            return
                        part = rr[rows[i][0]:rows[i][-1] + 1]
                        for c0 in range(0, len(part), 4096):
                            ref.intersect(rsc, part[c0:c0 + 4096], coherent=True)
                th = [threading.Thread(target=w) for _ in range(nth)]
                t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
                return time.perf_counter() - t0
            bt = 1e9
            for r in range(4):
                rr[:] = rays
                bt = min(bt, run_all())
            c = parity.compare_closest(ours, rr)
            res.update({"reference_mrays_per_s": n / bt / 1e6, "reference_threads": nth,
                        "parity": {k: c[k] for k in ("pass", "agreement", "hitmiss_disagree", "id_disagree_unexplained", "max_t_rel", "max_uv_abs")}})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
