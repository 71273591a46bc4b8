#!/usr/bin/env python3
"""Compile the UNMODIFIED reference (Embree 3.12.1 CPU path) into oracle/_ref/libembree3_ref.so.

TEST INFRASTRUCTURE ONLY.  The resulting library is the parity checker and the CPU baseline
(`bench.py --impl reference`, `cpu_baseline.kind == "reference"`); nothing in the product
(`embree-aarch64_b200/`) links, loads or calls it.

This is our own recipe, not the reference's build system: the reference's CMake project is not
run.  The sources are compiled where they lie under /root/reference with plain g++; the three
small configuration headers CMake would have generated from `kernels/config.h.in`,
`kernels/rtcore_config.h.in` and `kernels/hash.h.in` are written by this script into
oracle/_ref/gen/ (only #define lines: triangle, quad and instance geometry only, ray packets on, filter functions
on, ray masks off, backface culling off -- the reference's defaults, CMakeLists.txt:155-179).
All outputs go to oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

The source lists below restate kernels/CMakeLists.txt:24-230 (which files go into the lowest-ISA
library and which are re-compiled once per ISA) and common/*/CMakeLists.txt.  Tasking system is
the reference's INTERNAL scheduler (TBB is neither vendored nor installed here).

Usage: python oracle/build_ref.py [--ref /root/reference] [--isas sse42,avx,avx2,avx512skx] [-j N]
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

COMMON = {
    "sys": ["sysinfo", "alloc", "filename", "library", "thread", "string", "regression", "mutex",
            "condition", "barrier"],
    "math": ["constants"],
    "simd": ["sse"],
    "lexers": ["stringstream", "tokenstream"],
    "tasking": ["taskschedulerinternal"],
    "algorithms": ["parallel_for", "parallel_reduce", "parallel_prefix_sum", "parallel_for_for",
                   "parallel_for_for_prefix_sum", "parallel_partition", "parallel_sort",
                   "parallel_set", "parallel_map", "parallel_filter"],
}

# kernels/CMakeLists.txt:24-87,126-132  (lowest-ISA library)
LOWEST = """
common/device common/stat common/acceln common/accelset common/state common/rtcore
common/rtcore_builder common/scene common/alloc common/geometry common/scene_user_geometry
common/scene_instance common/scene_triangle_mesh common/scene_quad_mesh common/scene_curves
common/scene_line_segments common/scene_grid_mesh common/scene_points common/motion_derivative
subdiv/bezier_curve subdiv/bspline_curve subdiv/catmullrom_curve
geometry/primitive4 geometry/instance_intersector
geometry/curve_intersector_virtual geometry/curve_intersector_virtual2
geometry/curve_intersector_virtual_point geometry/curve_intersector_virtual_point2
geometry/curve_intersector_virtual_bezier_curve geometry/curve_intersector_virtual_bezier_curve2
geometry/curve_intersector_virtual_bspline_curve geometry/curve_intersector_virtual_bspline_curve2
geometry/curve_intersector_virtual_linear_curve geometry/curve_intersector_virtual_linear_curve2
geometry/curve_intersector_virtual_catmullrom_curve geometry/curve_intersector_virtual_catmullrom_curve2
geometry/curve_intersector_virtual_hermite_curve
builders/primrefgen
bvh/bvh bvh/bvh_statistics bvh/bvh4_factory bvh/bvh8_factory
bvh/bvh_collider bvh/bvh_rotate bvh/bvh_refit bvh/bvh_builder bvh/bvh_builder_hair
bvh/bvh_builder_hair_mb bvh/bvh_builder_morton bvh/bvh_builder_sah bvh/bvh_builder_sah_spatial
bvh/bvh_builder_sah_mb bvh/bvh_builder_twolevel
bvh/bvh_intersector1_bvh4
bvh/bvh_intersector_hybrid4_bvh4 bvh/bvh_intersector_stream_bvh4 bvh/bvh_intersector_stream_filters
""".split()

ISA_ORDER = ["sse42", "avx", "avx2", "avx512skx"]
ISA_FLAGS = {  # common/cmake/gnu.cmake:14-19
    "sse2": "-msse2",
    "sse42": "-msse4.2",
    "avx": "-mavx",
    "avx2": "-mf16c -mavx2 -mfma -mlzcnt -mbmi -mbmi2",
    "avx512skx": "-mavx512f -mavx512dq -mavx512cd -mavx512bw -mavx512vl -mf16c -mavx2 -mfma "
                 "-mlzcnt -mbmi -mbmi2 -mprefer-vector-width=256",
}


def isa_files(isa):
    """kernels/CMakeLists.txt:134-229 (macro embree_files)."""
    rank = {"sse42": 1, "avx": 2, "avx2": 3, "avx512skx": 5}[isa]
    f = """geometry/instance_intersector
geometry/curve_intersector_virtual geometry/curve_intersector_virtual2
geometry/curve_intersector_virtual_point geometry/curve_intersector_virtual_point2
geometry/curve_intersector_virtual_bezier_curve geometry/curve_intersector_virtual_bezier_curve2
geometry/curve_intersector_virtual_bspline_curve geometry/curve_intersector_virtual_bspline_curve2
geometry/curve_intersector_virtual_linear_curve geometry/curve_intersector_virtual_linear_curve2
geometry/curve_intersector_virtual_catmullrom_curve geometry/curve_intersector_virtual_catmullrom_curve2
geometry/curve_intersector_virtual_hermite_curve geometry/curve_intersector_virtual_hermite_curve2
bvh/bvh_intersector1_bvh4""".split()
    if isa == "avx":
        f += ["geometry/primitive8"]
    if isa in ("avx", "avx2", "avx512skx"):
        f += """common/scene_user_geometry common/scene_instance common/scene_triangle_mesh
common/scene_quad_mesh common/scene_curves common/scene_line_segments common/scene_grid_mesh
common/scene_points bvh/bvh_collider bvh/bvh_refit bvh/bvh_builder bvh/bvh_builder_hair
bvh/bvh_builder_hair_mb bvh/bvh_builder_sah bvh/bvh_builder_sah_spatial bvh/bvh_builder_sah_mb
bvh/bvh_builder_twolevel""".split()
    if isa in ("avx", "avx2"):
        f += ["bvh/bvh_builder_morton", "bvh/bvh_rotate", "builders/primrefgen"]
    if isa == "avx512skx":
        # the avx512skx objects reference avx512skx::createPrimRefArray* / *BuilderMorton*; the
        # reference's CMake leaves them undefined (never called, lazily bound); dlopen with
        # RTLD_NOW (ctypes) needs them, so they are compiled here as well
        f += ["builders/primrefgen", "bvh/bvh_builder_morton", "bvh/bvh_rotate"]
    if rank > 1:
        f += ["bvh/bvh_intersector1_bvh8"]
    if isa == "avx":
        f += ["bvh/bvh", "bvh/bvh_statistics"]
    f += ["bvh/bvh_intersector_hybrid4_bvh4", "bvh/bvh_intersector_stream_bvh4",
          "bvh/bvh_intersector_stream_filters"]
    if rank > 1:
        f += ["bvh/bvh_intersector_hybrid8_bvh4", "bvh/bvh_intersector_hybrid4_bvh8",
              "bvh/bvh_intersector_hybrid8_bvh8", "bvh/bvh_intersector_stream_bvh8"]
    if rank > 3:
        f += ["bvh/bvh_intersector_hybrid16_bvh8", "bvh/bvh_intersector_hybrid16_bvh4"]
    return f


CONFIG_H = """// written by oracle/build_ref.py (stands in for the CMake-configured kernels/config.h)
#define EMBREE_FILTER_FUNCTION
#define EMBREE_GEOMETRY_TRIANGLE
#define EMBREE_GEOMETRY_QUAD
#define EMBREE_GEOMETRY_INSTANCE
#define EMBREE_RAY_PACKETS
{stat}
#define EMBREE_CURVE_SELF_INTERSECTION_AVOIDANCE_FACTOR 2.0
#define IF_ENABLED_TRIS(x) x
#define IF_ENABLED_QUADS(x) x
#define IF_ENABLED_CURVES_OR_POINTS(x)
#define IF_ENABLED_CURVES(x)
#define IF_ENABLED_POINTS(x)
#define IF_ENABLED_SUBDIV(x)
#define IF_ENABLED_USER(x)
#define IF_ENABLED_INSTANCE(x) x
#define IF_ENABLED_GRIDS(x)
"""

RTCORE_CONFIG_H = """// written by oracle/build_ref.py (stands in for include/embree3/rtcore_config.h)
#pragma once
#define RTC_VERSION_MAJOR 3
#define RTC_VERSION_MINOR 12
#define RTC_VERSION_PATCH 1
#define RTC_VERSION 31201
#define RTC_VERSION_STRING "3.12.1"
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1
#define EMBREE_MIN_WIDTH 0
#define RTC_MIN_WIDTH EMBREE_MIN_WIDTH
#define RTC_NAMESPACE_BEGIN
#define RTC_NAMESPACE_END
#define RTC_NAMESPACE_USE
#if defined(__cplusplus)
#  define RTC_API_EXTERN_C extern "C"
#else
#  define RTC_API_EXTERN_C
#endif
#define RTC_API_IMPORT RTC_API_EXTERN_C
#define RTC_API_EXPORT RTC_API_EXTERN_C __attribute__ ((visibility ("default")))
#if defined(RTC_EXPORT_API)
#  define RTC_API RTC_API_EXPORT
#else
#  define RTC_API RTC_API_IMPORT
#endif
"""

EXPORT_MAP = "{\nglobal:\n  rtc*;\n  _ZN6embree13TaskScheduler*;\nlocal:\n  *;\n};\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--isas", default="sse42,avx,avx2,avx512skx")
    ap.add_argument("--stat-counters", action="store_true",
                    help="build a second library with EMBREE_STAT_COUNTERS (traversal step counts)")
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args()
    ref = os.path.abspath(args.ref)
    if not os.path.isdir(os.path.join(ref, "kernels")):
        print("reference tree not found at", ref, "- nothing to build", file=sys.stderr)
        return 0
    isas = [i for i in ISA_ORDER if i in args.isas.split(",")]
    tag = "stat" if args.stat_counters else "rel"
    libname = "libembree3_ref_stat.so" if args.stat_counters else "libembree3_ref.so"
    bdir = os.path.join(OUT, "build_" + tag)
    gen = os.path.join(OUT, "gen_" + tag)
    os.makedirs(os.path.join(gen, "kernels", "x"), exist_ok=True)
    os.makedirs(os.path.join(gen, "embree3"), exist_ok=True)
    os.makedirs(bdir, exist_ok=True)

    def write(path, text):
        if not os.path.exists(path) or open(path).read() != text:
            open(path, "w").write(text)

    # `#include "../config.h"` from kernels/common/default.h resolves through -I gen/kernels/x
    write(os.path.join(gen, "kernels", "config.h"),
          CONFIG_H.format(stat="#define EMBREE_STAT_COUNTERS" if args.stat_counters else ""))
    write(os.path.join(gen, "kernels", "hash.h"), '#define RTC_HASH "oracle_build_ref"\n')
    write(os.path.join(gen, "embree3", "rtcore_config.h"), RTCORE_CONFIG_H)
    write(os.path.join(gen, "export.map"), EXPORT_MAP)

    base = ("-std=c++11 -O3 -DNDEBUG -fPIC -fvisibility=hidden -fvisibility-inlines-hidden "
            "-fno-strict-aliasing -fno-tree-vectorize -fno-strict-overflow "
            "-fno-delete-null-pointer-checks -fwrapv -w -DTASKING_INTERNAL -DRTC_EXPORT_API "
            "-DEMBREE_TARGET_SSE2 " + " ".join("-DEMBREE_TARGET_" + i.upper() for i in isas) +
            f" -I{gen}/embree3 -I{gen}/kernels/x -I{gen}/kernels")

    rules = ["rule cxx\n  command = g++ $flags -MMD -MF $out.d -c $in -o $out\n  depfile = $out.d\n"
             "  deps = gcc\n  description = CXX $out\n",
             "rule link\n  command = g++ -shared -o $out @$out.rsp -Wl,--version-script=" +
             f"{gen}/export.map -lpthread -ldl\n  rspfile = $out.rsp\n  rspfile_content = $in\n"
             "  description = LINK $out\n"]
    objs = []

    def add(src, obj, flags):
        objs.append(obj)
        rules.append(f"build {obj}: cxx {src}\n  flags = {base} {flags}\n")

    for d, names in COMMON.items():
        for n in names:
            add(f"{ref}/common/{d}/{n}.cpp", f"{bdir}/common_{d}_{n}.o", ISA_FLAGS["sse2"])
    for f in LOWEST:
        add(f"{ref}/kernels/{f}.cpp", f"{bdir}/k_{f.replace('/', '_')}.o",
            ISA_FLAGS["sse2"] + " -DEMBREE_LOWEST_ISA")
    for isa in isas:
        for f in isa_files(isa):
            add(f"{ref}/kernels/{f}.cpp", f"{bdir}/k_{f.replace('/', '_')}.{isa}.o", ISA_FLAGS[isa])
    lib = os.path.join(OUT, libname)
    rules.append(f"build {lib}: link {' '.join(objs)}\n")
    rules.append(f"default {lib}\n")
    nf = os.path.join(bdir, "build.ninja")
    write(nf, "\n".join(rules))
    r = subprocess.call(["ninja", "-f", nf, "-j", str(args.j)])
    if r == 0:
        print("built", lib)
    return r


if __name__ == "__main__":
    sys.exit(main())
