"""The C + OpenMP ray generators (tinybvh_b200/hostgen/raygen.c) against their numpy definitions in tinybvh_b200/rays.py: the bench
builds its 2 x 67.1 M-ray workload with the former, the parity tests and golden vectors were made with the latter."""
import numpy as np
import pytest

from tinybvh_b200 import rays as R, scenes

pytestmark = pytest.mark.skipif(R._genlib() is None, reason="libtbvh_raygen.so not built")


def same(a, b):
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("spp", [16, 4, 2])
def test_primary_identical(spp):
    eye, view = R.SPONZA_EYES[1], R.SPONZA_VIEWS[1]
    want = R.primary_rays(eye, view, 64, 48, spp)
    got = np.empty(want.shape[0], R.RAY_DTYPE)
    R.primary_rays_into(got, eye, view, 64, 48, spp)
    assert same(got, want)
    # a slice of the set (what a rank of a sharded run generates)
    part = np.empty(1000, R.RAY_DTYPE)
    R.primary_rays_into(part, eye, view, 64, 48, spp, first=777)
    assert same(part, want[777:1777])


def test_shadow_and_diffuse_identical():
    v = scenes.procedural_scene(3000, 5)
    lo, hi = scenes.scene_bounds(v)
    eye, view = R.bounds_camera(lo, hi, "inside")
    prim = R.primary_rays(eye, view, 64, 64, 4)
    rng = np.random.default_rng(1)
    prim["t"] = np.where(rng.random(prim.shape[0]) < 0.7, rng.random(prim.shape[0]) * 60, 1e30).astype(np.float32)
    prim["prim"] = rng.integers(0, 3000, prim.shape[0])
    light = ((lo + hi) * 0.5).astype(np.float32)
    want = R.shadow_rays(prim, light, 1e-5)
    got = np.empty_like(want)
    R.shadow_rays_into(got, prim, light, 1e-5)
    assert same(got, want)
    # packed hits instead of the records' own
    hits = np.zeros((prim.shape[0], 4), np.float32)
    hits[:, 0], hits[:, 3] = prim["t"], prim["prim"].view(np.float32)
    bare = prim.copy()
    R.reset_hits(bare)
    R.shadow_rays_into(got, bare, light, 1e-5, hits=hits)
    assert same(got, want)
    want = R.diffuse_rays(prim, v)
    R.diffuse_rays_into(got, prim, v)
    bad = (got.view(np.uint8).reshape(-1, 128) != want.view(np.uint8).reshape(-1, 128)).any(1).sum()
    assert bad == 0, f"{bad} diffuse rays differ"
    R.diffuse_rays_into(got, bare, v, hits=hits)
    assert same(got, want)
    part = np.empty(500, R.RAY_DTYPE)
    R.diffuse_rays_into(part, prim[100:600], v, first=100)
    assert same(part, want[100:600])


def test_reset_hits():
    r = R.make_rays(np.zeros((10, 3), np.float32), np.ones((10, 3), np.float32))
    r["t"], r["u"], r["prim"] = 3, 0.5, 9
    R.reset_hits_fast(r)
    assert (r["t"] == np.float32(1e30)).all() and not r["u"].any() and not r["prim"].any()
