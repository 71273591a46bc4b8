"""Drop-in proof in C: programs written against the reference's API compile against include/embree3 and
link with -lembree3 from this package, unchanged.
  * the reference's own tutorials/minimal/minimal.cpp (compiled where /root/reference exists; the binary is kept under
    tests/_build/ -- git-ignored, shipped to the GPU box -- so the GPU test can run it there);
  * tests/dropin/stream_app.c: a C99 application (own code) driving rtcIntersect1M / rtcOccluded1M on malloc'ed streams plus
    the memory-monitor fault-injection check of verify.cpp:4564-4634."""
import os
import subprocess

import pytest

import cases

ROOT = cases.ROOT
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "embree-aarch64_b200", "lib")
OUT = os.path.join(ROOT, "tests", "_build")
REF_MINIMAL = "/root/reference/tutorials/minimal/minimal.cpp"


def _build(src, exe, compiler, std):
    os.makedirs(OUT, exist_ok=True)
    cmd = [compiler, std, "-O1", "-I" + INC, "-o", exe, src, "-L" + LIBDIR, "-lembree3", "-lm", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return exe


def test_reference_minimal_tutorial_compiles_and_links_unmodified():
    if not os.path.exists(REF_MINIMAL):
        pytest.skip("reference tree not mounted here")
    exe = _build(REF_MINIMAL, os.path.join(OUT, "minimal_ref"), "g++", "-std=c++11")
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    used = sorted({l.split()[-1].split("@")[0] for l in nm.splitlines() if " rtc" in l})
    assert "rtcIntersect1" in used and "rtcSetNewGeometryBuffer" in used and len(used) >= 12, used
    have = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libembree3.so")], capture_output=True, text=True).stdout
    for s in used:
        assert f" {s}\n" in have, s


def test_c_application_compiles_as_c99():
    _build(os.path.join(ROOT, "tests", "dropin", "stream_app.c"), os.path.join(OUT, "stream_app"), "gcc", "-std=c99")


@pytest.mark.gpu
def test_reference_minimal_tutorial_runs():
    exe = os.path.join(OUT, "minimal_ref")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/minimal_ref was not built (needs the reference tree at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert any("Found intersection on geometry 0, primitive 0 at tfar=1" in l for l in lines), out
    assert any("Did not find any intersection" in l for l in lines), out
    assert not any(l.startswith("error") for l in lines), out


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["", "gpu_builder=ploc"])
def test_c_application_streams_and_memory_monitor(cfg):
    exe = _build(os.path.join(ROOT, "tests", "dropin", "stream_app.c"), os.path.join(OUT, "stream_app"), "gcc", "-std=c99")
    r = subprocess.run([exe, cfg], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok done 0 0" in r.stdout and "FAIL" not in r.stdout, r.stdout + r.stderr
