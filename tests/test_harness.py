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


def test_speedtest_patch_applies_to_the_reference():
    """harness/speedtest_b200.patch is a patch of the reference's own tiny_bvh_speedtest.cpp (GPU section, :1071): it must apply
    cleanly to an untouched copy (only where the reference checkout exists - this container, not the GPU box)."""
    import shutil
    import tempfile
    ref = "/root/reference/tiny_bvh_speedtest.cpp"
    if not os.path.isfile(ref):
        pytest.skip("reference checkout not present")
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(ref, os.path.join(d, "tiny_bvh_speedtest.cpp"))
        r = subprocess.run(["patch", "-p1", "--dry-run", "-d", d, "-i", os.path.join(REPO, "harness", "speedtest_b200.patch")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        subprocess.check_call(["patch", "-s", "-p1", "-d", d, "-i", os.path.join(REPO, "harness", "speedtest_b200.patch")])
        src = open(os.path.join(d, "tiny_bvh_speedtest.cpp")).read()
        assert "#ifdef ENABLE_B200" in src and "tinybvh_b200::BVH8_CWBVH" in src


@pytest.mark.gpu
def test_patched_reference_speedtest_runs(gpu):
    """The literal drop-in: the reference's tiny_bvh_speedtest.cpp + harness/speedtest_b200.patch, built with -DENABLE_B200 (oracle/Makefile
    speedtest_patched), run as the reference runs it - from a directory holding ./testdata/cryteksponza.bin.  All of its CPU sections
    run too; no `!! Validation failed` line may appear, and the B200 section must report."""
    import tempfile
    exe = os.path.join(REPO, "oracle", "_ref", "tiny_bvh_speedtest_b200")
    scene = os.path.join(REPO, "data", "scenes", "cryteksponza.bin")
    if not (os.path.isfile(exe) and os.path.isfile(scene)):
        pytest.skip("patched speedtest binary or Sponza fixture not present")
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "testdata"))
        os.symlink(scene, os.path.join(d, "testdata", "cryteksponza.bin"))
        r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=900)
    print(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "BVH traversal speed - GPU (B200, tinybvh_b200)" in r.stdout
    assert "!! Validation" not in r.stdout.split("BVH traversal speed - GPU (B200, tinybvh_b200)")[1].split("BVH traversal speed - CPU multi-core")[0]
    assert "- BVH8_CWBVH  - primary:" in r.stdout and "(host buffers)" in r.stdout


@pytest.mark.gpu
def test_save_load_round_trips_with_the_reference(gpu):
    """BVH::Save / Load and BVH8_CWBVH::Save / Load (tiny_bvh.h:1747-1799, 5786-5820) between the engine and the compiled reference,
    both directions (harness/saveload_b200.cpp, built by `make -C oracle saveload` where the reference header exists)."""
    import tempfile
    exe = os.path.join(REPO, "oracle", "_ref", "saveload_b200")
    scene = os.path.join(REPO, "data", "scenes", "bunny.bin")
    if not (os.path.isfile(exe) and os.path.isfile(scene)):
        pytest.skip("saveload_b200 binary or bunny fixture not present")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([exe, scene, d], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all round trips ok" in r.stdout


def test_optimize_host_half_matches_the_reference_objects():
    """tinybvh_b200::BVH::Optimize = download -> FillReference -> the reference's own BVH::Optimize (tiny_bvh.h:3043) -> upload.  The part between the
    two transfers runs without a GPU: a tinybvh::BVH filled from plain arrays must behave under Optimize exactly like the object the reference's
    builder made (harness/optimize_host_check.cpp, built by `make -C oracle optimize_check` where the reference header exists)."""
    exe = os.path.join(REPO, "oracle", "_ref", "optimize_host_check")
    scene = os.path.join(REPO, "data", "scenes", "bunny.bin")
    if not (os.path.isfile(exe) and os.path.isfile(scene)):
        pytest.skip("optimize_host_check binary or bunny fixture not present")
    r = subprocess.run([exe, scene, "3"], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "host half of Optimize ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
