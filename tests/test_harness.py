"""The C++ side of the boundary: the shim header compiles against the C-ABI (CPU), and on the GPU box the reference-free
example and the re-hosted speedtest GPU section run and validate."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_header_compiles_and_links():
    from tinybvh_b200 import build
    build.build()
    out = os.path.join(REPO, "harness", "minimal_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "harness", "minimal_b200.cpp"),
                           "-L" + os.path.join(REPO, "tinybvh_b200"), "-ltinybvh_b200", "-Wl,-rpath,$ORIGIN/../tinybvh_b200", "-o", out])
    assert os.path.isfile(out)


def test_shim_mirrors_reference_names():
    hpp = open(os.path.join(REPO, "include", "tinybvh_b200.hpp")).read()
    for name in ("class BVH ", "class BVH_GPU", "class BVH8_CWBVH", "void Build(", "BuildHQ(", "ConvertFrom(", "Intersect(", "IsOccluded(", "c_trav", "usedNodes"):
        assert name in hpp, name


@pytest.mark.gpu
def test_minimal_example_runs(gpu):
    exe = os.path.join(REPO, "harness", "minimal_b200")
    if not os.path.isfile(exe):
        pytest.skip("harness/minimal_b200 not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rays hit" in r.stdout


@pytest.mark.gpu
def test_speedtest_gpu_section_drop_in(gpu):
    exe = os.path.join(REPO, "oracle", "_ref", "speedtest_b200")
    scene = os.path.join(REPO, "data", "scenes", "cryteksponza.bin")
    if not (os.path.isfile(exe) and os.path.isfile(scene)):
        pytest.skip("speedtest_b200 binary or Sponza fixture not present")
    r = subprocess.run([exe, scene, "320", "240"], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 prim mismatches, 0 t-bit mismatches" in r.stdout
