"""How evenly does the bench batch split over N GPUs?  One GPU plays every rank in turn.
The seed-0 diffuse + shadow streams of the 10 M-triangle scene are generated in 64-row bands; a rank's shard is a set of bands
(contiguous: the N-th part of the frame; cyclic:B = blocks of B bands dealt round-robin).  Prints the device time of every shard
and max / mean -- the strong-scaling efficiency the max-over-ranks rule can reach at best.
usage: python tools/shard_balance.py [--world 8] [--workload c3]"""
import argparse, importlib, json, sys
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("embree-aarch64_b200")
fx, rt = pkg.fixtures, pkg.rtcore
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--workload", default="c3")
ap.add_argument("--band", type=int, default=64)
ap.add_argument("--lib", default=None)
args = ap.parse_args()
lib = rt.RTCore(args.lib) if args.lib else rt.RTCore()
dev = lib.new_device("async=1")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
lib.lib.rtcxSetDeviceStream(dev, st.cuda_stream)
sc, keep = lib.build_scene(dev, fx.scene_c3(1.0) if args.workload == "c3" else fx.scene_c2(1.0))
nb = 4096 // args.band
D, S = [], []
for b in range(nb):
    prim = fx.primary_rays(4096, 4096, rows=(b * args.band, (b + 1) * args.band), **fx.C2_CAMERA)
    lib.intersect(sc, prim, coherent=True)
    D.append(fx.diffuse_rays(prim)); S.append(fx.shadow_rays(prim))


def shard_time(bands):
    d = np.concatenate([D[b] for b in bands]); s = np.concatenate([S[b] for b in bands])
    p_d = torch.from_numpy(d.view(np.uint8).reshape(len(d), 80)).cuda(); p_s = torch.from_numpy(s.view(np.uint8).reshape(len(s), 48)).cuda()
    w_d, w_s = p_d.clone(), p_s.clone()
    best = [1e9, 1e9]
    for rep in range(4):
        w_d.copy_(p_d); w_s.copy_(p_s); torch.cuda.synchronize()
        e = [torch.cuda.Event(True) for _ in range(3)]
        e[0].record(); lib.intersect_ptr(sc, w_d.data_ptr(), len(d), 80)
        e[1].record(); lib.occluded_ptr(sc, w_s.data_ptr(), len(s), 48)
        e[2].record(); torch.cuda.synchronize()
        if rep:
            best = [min(best[0], e[0].elapsed_time(e[1])), min(best[1], e[1].elapsed_time(e[2]))]
    return best[0], best[1], len(d) + len(s)


W = args.world
whole = shard_time(range(nb))
print(json.dumps({"shards": "whole frame", "closest_ms": whole[0], "occluded_ms": whole[1], "rays": whole[2]}), flush=True)
plans = {"contiguous": [list(range(r * nb // W, (r + 1) * nb // W)) for r in range(W)]}
for blk in (1, 2, 4):
    if nb // (W * blk) >= 1:
        plans[f"cyclic:{blk}"] = [[b for b in range(nb) if (b // blk) % W == r] for r in range(W)]
for name, plan in plans.items():
    t = [shard_time(p) for p in plan]
    tot = [a + b for a, b, _ in t]
    print(json.dumps({"shards": name, "world": W, "band_rows": args.band, "ms_per_rank": [round(x, 3) for x in tot], "rays_per_rank": [n for _, _, n in t],
                      "max_ms": max(tot), "mean_ms": sum(tot) / W, "balance": (sum(tot) / W) / max(tot),
                      "speedup_vs_whole": (whole[0] + whole[1]) / max(tot)}), flush=True)
