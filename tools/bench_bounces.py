"""BASELINE.json configs[2]: 10 M-triangle scene, 2-bounce path-tracer style secondary streams.  A 4096x4096 primary pass
(coherent), then per bounce a cosine-weighted diffuse stream from every hit (closest-hit, incoherent) followed by a shadow
stream towards a point light (occlusion); only rays that hit continue.  Each stream is timed device-resident with CUDA
events; the reference library traces a 1 M-ray sample of every stream on the host cores, and the two answers are compared.
usage: python tools/bench_bounces.py [--workload c3|c2] [--bounces 2] [--no-reference]"""
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


def ref_trace(ref, sc, rays, occluded, nth):
    parts = np.array_split(np.arange(len(rays)), nth * 4)
    nxt = [0]; lock = threading.Lock()

    def w():
        while True:
            with lock:
                i = nxt[0]; nxt[0] += 1
            if i >= len(parts):
                return
            p = rays[parts[i][0]:parts[i][-1] + 1]
            for c0 in range(0, len(p), 4096):
                (ref.occluded if occluded else ref.intersect)(sc, p[c0:c0 + 4096])
    th = [threading.Thread(target=w) for _ in range(nth)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--bounces", type=int, default=2)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    import torch
    meshes = fx.scene_c3(1.0) if a.workload == "c3" else fx.scene_c2(1.0)
    lib = rt.RTCore()
    dev = lib.new_device("async=1")
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    lib.lib.rtcxSetDeviceStream(dev, st.cuda_stream)
    sc, keep = lib.build_scene(dev, meshes)
    ref = rsc = None
    nth = os.cpu_count() or 1
    if not a.no_reference:
        from oracle.rq_oracle import REF_LIB
        if os.path.exists(REF_LIB):
            ref = rt.RTCore(REF_LIB)
            rdev = ref.new_device(f"threads={nth}")
            rsc, rkeep = ref.build_scene(rdev, meshes)

    def timed(stream_np, occluded):
        rec = 48 if occluded else 80
        n = len(stream_np)
        pristine = torch.from_numpy(stream_np.view(np.uint8).reshape(n, rec)).cuda()
        work = torch.empty_like(pristine)
        ts = []
        for r in range(4):
            work.copy_(pristine)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            (lib.occluded_ptr if occluded else lib.intersect_ptr)(sc, work.data_ptr(), n, rec)
            e1.record(); torch.cuda.synchronize()
            if r:
                ts.append(e0.elapsed_time(e1))
        out = work.cpu().numpy().reshape(-1).view(rt.RAY_DTYPE if occluded else rt.RAYHIT_DTYPE)
        info = {"rays": n, "ours_ms": min(ts), "ours_mrays_per_s": n / min(ts) / 1e3}
        if ref is not None:
            m = min(n, 1 << 20)
            lo = (n - m) // 2
            sample = stream_np[lo:lo + m].copy()
            dt = ref_trace(ref, rsc, sample, occluded, nth)
            c = parity.compare_occluded(out[lo:lo + m], sample) if occluded else parity.compare_closest(out[lo:lo + m], sample)
            info.update({"reference_mrays_per_s": m / dt / 1e6, "reference_sample": m, "parity_pass": c["pass"], "agreement": c["agreement"]})
        return out, info

    res = {"workload": f"configs[2]: {fx.num_tris(meshes)} triangles, 4096x4096 primary pass, {a.bounces} diffuse bounces, each followed by a shadow stream",
           "reference_threads": nth, "streams": []}
    parts = []
    for b in range(8):
        prim = fx.primary_rays(4096, 4096, rows=(b * 512, (b + 1) * 512), **fx.C2_CAMERA)
        lib.intersect(sc, prim, coherent=True)
        parts.append(prim)
    cur = np.concatenate(parts); del parts
    tot_rays, tot_ms = 0, 0.0
    for bounce in range(1, a.bounces + 1):
        diffuse = fx.diffuse_rays(cur, sample_id=bounce - 1)
        shadow = fx.shadow_rays(cur)
        del cur
        sh_out, si = timed(shadow, True); si["stream"] = f"shadow after bounce {bounce - 1}"
        d_out, di = timed(diffuse, False); di["stream"] = f"diffuse bounce {bounce}"
        di["hit_fraction"] = float((d_out["geomID"] != 0xFFFFFFFF).mean()); si["occluded_fraction"] = float(np.isneginf(sh_out["tfar"]).mean())
        res["streams"] += [si, di]
        tot_rays += si["rays"] + di["rays"]; tot_ms += si["ours_ms"] + di["ours_ms"]
        cur = d_out
        del shadow, diffuse, sh_out
    res["aggregate_mrays_per_s"] = tot_rays / tot_ms / 1e3
    res["total_rays"] = tot_rays
    print(json.dumps(res))


if __name__ == "__main__":
    main()
