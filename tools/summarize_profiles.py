#!/usr/bin/env python3
"""Turns the raw artefacts of tools/gpu_round.sh (gpurun_out/<tag>_*) into the small, tracked summaries
under profiles/:  <tag>_launches_summary.txt (per-kernel time shares of the ncu launch list),
<tag>_ncu_trace.txt (key counters of the --set full capture of the traversal kernels) and
ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).
Usage: python tools/summarize_profiles.py <tag>            (runs here, no GPU needed; needs ncu for the .ncu-rep)"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1]
workloads = sys.argv[2:] or ["c3", "c2"]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    return re.sub(r"\(.*", "", name)


def launches():
    path = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        k = short(r[4]); ns = float(r[14])
        c = per.setdefault(k, [0, 0.0, r[7], r[8]]); c[0] += 1; c[1] += ns
    total = sum(v[1] for v in per.values())
    with open(os.path.join(P, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, first {len(rows)} launches of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes. Columns: kernel, launches, total ms, share, block, grid(last)\n")
        for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} {v[0]:5d} {v[1] / 1e6:10.3f} ms {100 * v[1] / total:6.2f} %  {v[2]} {v[3]}\n")
        f.write(f"{'TOTAL':60s} {len(rows):5d} {total / 1e6:10.3f} ms\n")
    print(open(os.path.join(P, f"{tag}_launches_summary.txt")).read())


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * m.get(unit, 1)


def ncu(suffix="trace", what="configs[1] streams"):
    rep = os.path.join(G, f"{tag}_{suffix}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(os.path.join(P, f"{tag}_ncu_{suffix}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch each (tools/profile_trace.py, {what}); source: gpurun_out/{tag}_{suffix}.ncu-rep\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")]) + r[hdr.index("Kernel Name")][r[hdr.index("Kernel Name")].find("<", 12):][:40]
            f.write(f"\n== {r[hdr.index('Kernel Name')][:110]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:90s} {r[i]:>18s} {units[i]}\n")
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            kind = "occluded" if re.search(r"k_trace<\(?(bool\))?1|k_trace<true", r[hdr.index("Kernel Name")]) else "closest"
            traffic[f"dram_bytes_per_launch_{kind}"] = rd + wr
            if kind == "closest":
                for key, out in (("lts__t_sector_hit_rate.pct", "lts_hit_rate_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_rate_pct"),
                                 ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"), ("gpu__time_duration.sum", "duration"),
                                 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
                                 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct")):
                    if key in hdr:
                        traffic[out] = r[hdr.index(key)] + ("" if out != "duration" else " " + units[hdr.index(key)])
    print(open(os.path.join(P, f"{tag}_ncu_{suffix}.txt")).read())
    return traffic


launches()
import bench  # noqa: E402  (kernel_source_hash)
path = os.path.join(P, "ncu_traffic.json")
try:
    allt = json.load(open(path))
    if "dram_bytes_per_launch_closest" in allt:          # round-1 layout
        allt = {}
except Exception:
    allt = {}
for w in workloads:
    names = {"c3": "configs[2]-[3] scene (10 M triangles), 16.7 M diffuse + shadow rays", "c2": "configs[1] scene (1.0 M triangles), 16.7 M diffuse + shadow rays"}
    t = ncu(f"trace_{w}", names.get(w, w))
    if t:
        meta = {}
        try:
            meta = json.load(open(os.path.join(G, f"{tag}_trace_{w}.meta.json")))
        except Exception:
            pass
        t.update({"capture": f"profiles/{tag}_ncu_trace_{w}.txt", "kernel_source_hash": meta.get("kernel_source_hash", bench.kernel_source_hash()),
                  "git": meta.get("git"), "rays_per_launch": meta.get("rays_per_launch"),
                  "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one closest-hit / one occlusion launch (tools/profile_trace.py)"})
        allt[w] = t
json.dump(allt, open(path, "w"), indent=1)
for f in ("bench.json", "bench_reference.json", "pytest_gpu.log", "c3.log"):
    src = os.path.join(G, f"{tag}_{f}")
    if os.path.exists(src):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(src).read())
