"""Parity at the size BASELINE.json's configs[2] names: Bistro exterior (2,837,209 triangles).  Skipped when the fixture is
not present on the box (it is 136 MB and is only pushed for dedicated sessions; data/scenes is git-ignored)."""
import numpy as np
import pytest

from oracle import refpy
from tinybvh_b200 import api, rays as R, scenes
from tests import util

pytestmark = [pytest.mark.gpu]


def bistro():
    try:
        return scenes.load_scene("bistro", allow_synthetic=False)[0]
    except FileNotFoundError:
        pytest.skip("Bistro fixture not on this box")


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
def test_bistro_build_identical_and_incoherent_parity(gpu):
    v = bistro()
    o = refpy.RefBVH(v, mode=0, threaded=False)
    e = api.BVH().Build(v)
    nodes, idx = e.download()
    assert nodes.shape[0] == o.used_nodes
    assert np.array_equal(nodes.view(np.uint32), o.nodes.view(np.uint32)), "GPU-built Bistro tree differs from BVH::Build"
    assert np.array_equal(idx, o.prim_idx)
    lo, hi = scenes.scene_bounds(v)
    eye, view = R.bounds_camera(lo, hi, "inside")
    want = R.primary_rays(eye, view, 512, 512, 4)
    got = want.copy()
    o.intersect(want), e.Intersect(got)
    assert util.compare_hits(got, want) == {"prim": 0, "t": 0, "u": 0, "v": 0}
    d = util.derived_sets(want, v, (lo, hi))
    a, b = d["diffuse"].copy(), d["diffuse"].copy()
    o.intersect(a), e.Intersect(b)
    assert util.compare_hits(b, a) == {"prim": 0, "t": 0, "u": 0, "v": 0}, "incoherent rays: not bit-exact on the oracle's own tree"
    assert np.array_equal(e.IsOccluded(d["shadow"]), o.occluded(d["shadow"]))
    print("bistro: build", e.info().build_ms, "ms", v.shape[0] // 3 / e.info().build_ms / 1e3, "Mtris/s, depth", e.info().max_depth)


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
def test_bistro_cwbvh_conversion_and_traversal(gpu):
    v = bistro()
    cw = refpy.RefCWBVH(v, mode=0)   # BVH8_CWBVH::Build itself
    e = api.BVH8_CWBVH().Build(v)
    nodes, tris = e.download()
    assert nodes.shape == cw.nodes.shape and np.array_equal(nodes.view(np.uint32), cw.nodes.view(np.uint32)), "bvh8Data differs"
    assert np.array_equal(tris.view(np.uint32), cw.tris.view(np.uint32)), "bvh8Tris differs"
    lo, hi = scenes.scene_bounds(v)
    eye, view = R.bounds_camera(lo, hi, "inside")
    prim = R.primary_rays(eye, view, 256, 256, 4)
    tr = prim.copy()
    cw.intersect(tr)
    d = R.diffuse_rays(tr, v)
    a, b = d.copy(), d.copy()
    cw.intersect(a), e.Intersect(b)
    assert util.compare_hits(b, a) == {"prim": 0, "t": 0, "u": 0, "v": 0}


def test_bistro_build_hq_identical(gpu):
    """BVH::BuildHQ at Bistro's size: 2.8 M triangles, 1.5x index slack, thousands of spatial splits per level."""
    from oracle import portpy
    v = bistro()
    nodes_want, idx_want, ic = portpy.build_hq(v)
    e = api.BVH().BuildHQ(v)
    nodes, idx = e.download()
    assert e.info().idx_count == ic and nodes.shape[0] == nodes_want.shape[0]
    assert np.array_equal(nodes.view(np.uint32), nodes_want.view(np.uint32)), "GPU-built Bistro SBVH differs from BVH::BuildHQ"
    assert np.array_equal(idx[: idx_want.shape[0]], idx_want) and not idx[idx_want.shape[0]:].any()
    print("bistro: BuildHQ", e.info().build_ms, "ms, nodes", nodes.shape[0], "refs", idx_want.shape[0], "depth", e.info().max_depth)


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
def test_bistro_cwbvh_build_hq(gpu):
    """BVH8_CWBVH::BuildHQ (:5859) at Bistro's size: SBVH -> SplitLeafs(3) -> 8-wide collapse -> CWBVH, all on the GPU."""
    v = bistro()
    cw = refpy.RefCWBVH(v, mode=1)
    e = api.BVH8_CWBVH().BuildHQ(v)
    nodes, tris = e.download()
    assert nodes.shape == cw.nodes.shape and np.array_equal(nodes.view(np.uint32), cw.nodes.view(np.uint32)), "bvh8Data differs"
    used = int(cw.source_bvh().nodes["triCount"].sum()) * 3
    assert np.array_equal(tris[:used].view(np.uint32), cw.tris[:used].view(np.uint32)), "bvh8Tris differs"


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
def test_lucy_dragon_x29_build_and_traversal_parity(gpu):
    """BASELINE.json configs[4] at its size: lucy + xyzrgb_dragon replicated 29x (10,145,708 triangles).  The GPU-built tree is
    the reference's byte for byte; camera, shadow and bounce rays are bit-identical to BVH::Intersect / IsOccluded; the CWBVH layout is
    bit-identical to the reference's own walk of its own conversion."""
    try:
        v, label = scenes.load_scene("lucy_dragon_x29", allow_synthetic=False)
    except FileNotFoundError:
        pytest.skip("lucy / dragon fixtures not on this box")
    o = refpy.RefBVH(v, mode=0, threaded=False)
    e = api.BVH().Build(v)
    nodes, idx = e.download()
    assert nodes.shape[0] == o.used_nodes
    assert np.array_equal(nodes.view(np.uint32), o.nodes.view(np.uint32)), "GPU-built 10M-triangle tree differs from BVH::Build"
    assert np.array_equal(idx, o.prim_idx)
    lo, hi = scenes.scene_bounds(v)
    want = R.primary_rays(*R.bounds_camera(lo, hi, "outside"), 512, 512, 4)
    got = want.copy()
    o.intersect(want), e.Intersect(got)
    assert util.compare_hits(got, want) == {"prim": 0, "t": 0, "u": 0, "v": 0}
    d = util.derived_sets(want, v, (lo, hi))
    a, b = d["diffuse"].copy(), d["diffuse"].copy()
    o.intersect(a), e.Intersect(b)
    assert util.compare_hits(b, a) == {"prim": 0, "t": 0, "u": 0, "v": 0}
    assert np.array_equal(e.IsOccluded(d["shadow"]), o.occluded(d["shadow"]))
    print("lucy_dragon_x29: build", e.info().build_ms, "ms", v.shape[0] // 3 / e.info().build_ms / 1e3, "Mtris/s, depth", e.info().max_depth)
    # CWBVH over the same tree: the chain of BVH8_CWBVH::Build applied to the scalar builder's tree (refpy mode 2)
    cw = refpy.RefCWBVH(v, mode=2)
    api.check(api._lib.lib().tbvh_convert(e.h, api.LAYOUT_CWBVH))
    e.layout = api.LAYOUT_CWBVH
    i = e.info()
    d8 = np.zeros((i.used_blocks, 4), np.float32)
    t8 = np.zeros((i.cwbvh_tri_count * 3, 4), np.float32)
    api.check(api._lib.lib().tbvh_download_cwbvh(e.h, d8.ctypes.data, t8.ctypes.data, api.HOST))
    assert d8.shape == cw.nodes.shape and np.array_equal(d8.view(np.uint32), cw.nodes.view(np.uint32)), "bvh8Data differs"
    assert np.array_equal(t8.view(np.uint32), cw.tris.view(np.uint32)), "bvh8Tris differs"
    a, b = d["diffuse"].copy(), d["diffuse"].copy()
    cw.intersect(a), e.Intersect(b)
    assert util.compare_hits(b, a) == {"prim": 0, "t": 0, "u": 0, "v": 0}
