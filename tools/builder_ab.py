"""Same-box A/B of builder configurations: build statistics (SAH by the reference formula, phases), traversal work per ray
(instrumented kernel) and Mrays/s of a diffuse closest-hit stream + its shadow stream on BASELINE scenes.
usage: python tools/builder_ab.py [--scenes c2,c3] [--cfgs "gpu_builder=ploc;gpu_builder=sah;..."] [--rows 1024] [--check]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("embree-aarch64_b200")
rt, fx = pkg.rtcore, pkg.fixtures


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="c2,c3")
    ap.add_argument("--cfgs", default="gpu_builder=ploc;gpu_builder=lbvh;gpu_builder=sah;gpu_builder=sah,treelet=256")
    ap.add_argument("--rows", type=int, default=1024)
    ap.add_argument("--check", action="store_true", help="structural check + independent SAH of the exported image (oracle/rq_image.py)")
    a = ap.parse_args()
    import torch
    lib = rt.RTCore(os.environ.get("RQ_B200_LIB")) if os.environ.get("RQ_B200_LIB") else rt.RTCore()
    for scene in a.scenes.split(","):
        if scene.startswith("soup"):
            meshes = fx.random_soup(int(float(scene[4:] or 10) * 1e6))
        else:
            meshes = fx.scene_c3(1.0) if scene == "c3" else fx.scene_c3(float(scene[2:])) if scene.startswith("c3x") else fx.scene_c2(1.0)
        ntris = fx.num_tris(meshes)
        rays_d = rays_s = None
        for cfg in a.cfgs.split(";"):
            dev = lib.new_device(cfg)
            sc, keep = lib.build_scene(dev, meshes)
            err = lib.lib.rtcGetDeviceError(dev)
            if err:
                print(json.dumps({"scene": scene, "cfg": cfg, "error": err}), flush=True)
                continue
            for g in range(len(meshes)):
                lib.lib.rtcCommitGeometry(lib.lib.rtcGetGeometry(sc, g))
            t0 = time.perf_counter(); lib.lib.rtcCommitScene(sc); wall = (time.perf_counter() - t0) * 1e3
            st = lib.build_stats(sc)
            if rays_d is None:                                   # the same streams for every configuration of a scene
                if scene.startswith("soup"):
                    rs = fx.RandomSampler(np.arange(1 << 22), 5)
                    o = np.stack([rs.get_float() for _ in range(3)], 1).astype(np.float32)
                    d = np.stack([rs.get_float() * 2 - 1 for _ in range(3)], 1).astype(np.float32)
                    rays_d = fx._set(rt.new_rays(1 << 22), o, d, 0.0, np.inf)
                    rays_s = fx.to_ray(rays_d); rays_s["tfar"] = 0.5
                else:
                    r0 = (4096 - a.rows) // 2
                    prim = fx.primary_rays(4096, 4096, rows=(r0, r0 + a.rows), **fx.C2_CAMERA)
                    lib.intersect(sc, prim, coherent=True)
                    rays_d, rays_s = fx.diffuse_rays(prim), fx.shadow_rays(prim)
            nd, ns = len(rays_d), len(rays_s)
            p_d = torch.from_numpy(rays_d.view(np.uint8).reshape(nd, 80).copy()).cuda()
            p_s = torch.from_numpy(rays_s.view(np.uint8).reshape(ns, 48).copy()).cuda()
            w_d, w_s = p_d.clone(), p_s.clone()
            cc = lib.intersect_counted(sc, w_d.data_ptr(), nd, 80)
            co = lib.intersect_counted(sc, w_s.data_ptr(), ns, 48, occluded=True)
            tc, to = [], []
            for _ in range(5):
                w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
                e = [torch.cuda.Event(True) for _ in range(3)]
                s = torch.cuda.current_stream()
                lib.lib.rtcxSetDeviceStream(dev, s.cuda_stream)
                e[0].record(); lib.intersect_ptr(sc, w_d.data_ptr(), nd, 80); e[1].record(); lib.occluded_ptr(sc, w_s.data_ptr(), ns, 48); e[2].record()
                torch.cuda.synchronize()
                tc.append(e[0].elapsed_time(e[1])); to.append(e[1].elapsed_time(e[2]))
            out = {"scene": scene, "tris": ntris, "cfg": cfg, "sah": st["sah"], "sahInner": st["sahInner"], "sahLeafTris": st["sahLeafTris"],
                   "sah4": st["sahInner"] + st["sahLeafTris"] / 4, "nodes": st["numNodes"], "leaves": st["numLeaves"], "depth": st["depth"],
                   "treelets": st["numTreelets"], "ploc_iters": st["builderIterations"], "build_ms": st["msTotal"], "wall_ms": wall,
                   "phases": {k: round(st[k], 3) for k in ("msPrims", "msSort", "msHierarchy", "msRefit", "msEmit")},
                   "closest_nodes": cc["nodes"] / nd, "closest_tris": cc["tris"] / nd, "late": cc["lateNodes"] / max(cc["nodes"], 1),
                   "empty": cc["emptyNodes"] / max(cc["nodes"], 1), "occ_nodes": co["nodes"] / max(co["rays"], 1), "occ_tris": co["tris"] / max(co["rays"], 1),
                   "closest_mrays": nd / min(tc) / 1e3, "occluded_mrays": ns / min(to) / 1e3, "hits": int(cc["hits"])}
            if a.check:
                from oracle import rq_image
                img = rq_image.fetch(lib, sc)
                img.check_structure()
                out["sah_independent"] = img.sah()[0]
            print(json.dumps(out), flush=True)
            lib.lib.rtcReleaseScene(sc); lib.lib.rtcReleaseDevice(dev)
            del p_d, p_s, w_d, w_s
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
