"""Work per ray of the REFERENCE (its own STAT counters, kernels/common/stat.h:36-100) on our bench streams.
Runs the stats-enabled build (oracle/build_ref.py --stat-counters) here on the CPU; the counters are global and
printed at process exit, so the diffuse stream's share = run('both') - run('primary').  Test/measurement tooling only."""
import importlib, re, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
LIB = "oracle/_ref/libembree3_ref_stat.so"


def child(workload, mode, rows):
    pkg = importlib.import_module("embree-aarch64_b200")
    fx, rt = pkg.fixtures, pkg.rtcore
    ref = rt.RTCore(LIB)
    dev = ref.new_device("threads=8")
    meshes = fx.scene_c3(1.0) if workload == "c3" else fx.scene_c2(1.0)
    sc, keep = ref.build_scene(dev, meshes)
    prim = fx.primary_rays(4096, 4096, rows=rows, **fx.C2_CAMERA)
    ref.intersect(sc, prim, coherent=True)
    if mode == "both":
        d = fx.diffuse_rays(prim)
        ref.intersect(sc, d)
        print("DIFFUSE_RAYS", len(d), "HITS", int((d["geomID"] != 0xFFFFFFFF).sum()), flush=True)


def parse(out):
    vals = {}
    sect = None
    for line in out.splitlines():
        m = re.match(r"\s*(\w[\w.]*)\s*=\s*([0-9.eE+-]+)", line)
        if "normal" in line.lower() and "=" not in line: sect = "normal"
        if m: vals.setdefault(m.group(1), float(m.group(2)))
    return vals


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2], sys.argv[3], (1792, 1792 + int(sys.argv[4])))
        sys.exit(0)
    workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
    rows = sys.argv[2] if len(sys.argv) > 2 else "128"
    outs = {}
    for mode in ("primary", "both"):
        r = subprocess.run([sys.executable, __file__, "child", workload, mode, rows], capture_output=True, text=True)
        outs[mode] = r.stdout + r.stderr
        open(f"/tmp/ref_stats_{workload}_{mode}.log", "w").write(outs[mode])
    print(outs["both"][-3000:])
