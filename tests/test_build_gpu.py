"""GPU parity of the binned-SAH builder: the tree built by the CUDA kernels (tbvh_build through the C-ABI) must be the
reference's BVH::Build tree byte for byte - node array (bounds, leftFirst, triCount, numbering) and primIdx."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import golden_util as G
from tests import util

pytestmark = pytest.mark.gpu


def assert_same_tree(e, nodes_want, idx_want, label=""):
    nodes, idx = e.download()
    assert nodes.shape[0] == nodes_want.shape[0], f"{label}: usedNodes {nodes.shape[0]} != {nodes_want.shape[0]}"
    a, b = nodes.view(np.uint32).reshape(-1, 8), np.ascontiguousarray(nodes_want).view(np.uint32).reshape(-1, 8)
    bad = np.nonzero((a != b).any(1))[0]
    assert bad.size == 0, f"{label}: {bad.size} nodes differ, first {bad[:5]}: got {a[bad[0]]} want {b[bad[0]]}"
    assert np.array_equal(idx, idx_want), f"{label}: primIdx differs at {np.nonzero(idx != idx_want)[0][:8]}"


@pytest.mark.parametrize("path", G.golden_files(), ids=lambda p: p.split("/")[-1])
def test_build_matches_golden_tree(gpu, path):
    g = G.load(path)
    e = api.BVH().Build(g["verts"])
    assert_same_tree(e, g["nodes"].view(np.uint8).view(api.NODE32).reshape(-1), g["prim_idx"], path)


@pytest.mark.parametrize("ntris,seed", [(1, 1), (2, 2), (3, 3), (31, 4), (256, 5), (257, 6), (300, 7), (1000, 8), (5000, 9), (70000, 10), (400000, 11)])
def test_build_matches_oracle_on_seeded_scenes(gpu, ntris, seed):
    v = scenes.procedural_scene(ntris, seed)
    o = util.oracle_bvh(v)
    e = api.BVH().Build(v)
    assert_same_tree(e, o.nodes, o.prim_idx, f"{ntris} tris")
    assert e.info().build_ms > 0


def test_build_degenerate_inputs(gpu):
    rng = np.random.default_rng(3)
    # all triangles identical (every split candidate has an empty side -> one big leaf)
    one = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    same = np.tile(one, (700, 1))
    # flat soup: zero extent on y
    flat = scenes.procedural_scene(3000, 12)
    flat[:, 1] = 1.5
    # clustered duplicates with exact centroid ties
    base = scenes.procedural_scene(900, 13)
    dup = np.concatenate([base, base, base[:300]])
    for name, v in (("identical", same), ("flat", flat), ("duplicates", dup)):
        o = util.oracle_bvh(v)
        e = api.BVH().Build(v)
        assert_same_tree(e, o.nodes, o.prim_idx, name)


@pytest.mark.parametrize("scene", ["bunny", "sponza"])
def test_build_fixtures_and_trace(gpu, scene):
    v, label = scenes.load_scene(scene)
    o = util.oracle_bvh(v)
    e = api.BVH().Build(v)
    assert_same_tree(e, o.nodes, o.prim_idx, label)
    lo, hi = scenes.scene_bounds(v)
    eye, view = (R.SPONZA_EYES[0], R.SPONZA_VIEWS[0]) if scene == "sponza" else R.bounds_camera(lo, hi, "outside")
    want = R.primary_rays(eye, view, 128, 128, 4)
    got = want.copy()
    o.intersect(want), e.Intersect(got)
    assert util.compare_hits(got, want) == {"prim": 0, "t": 0, "u": 0, "v": 0}


def test_rebuild_is_deterministic(gpu):
    v = scenes.procedural_scene(30000, 14)
    a = api.BVH().Build(v).download()
    b = api.BVH().Build(v).download()
    assert np.array_equal(a[0].view(np.uint8), b[0].view(np.uint8)) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("ntris,seed", [(1, 81), (2, 82), (129, 83), (1000, 84), (70000, 85), (300000, 86)])
def test_build_avx_flavour_matches_reference(gpu, ntris, seed):
    """BVH::BuildAVX (what BuildDefault runs on x86): same tree as the reference's, byte for byte."""
    from oracle import portpy, refpy
    v = scenes.procedural_scene(ntris, seed)
    o = refpy.RefBVH(v, mode=1, threaded=False) if refpy.available() else portpy.PortBVH(v, avx=True)
    e = api.BVH().BuildAVX(v)
    assert_same_tree(e, o.nodes, o.prim_idx, f"BuildAVX {ntris} tris")


@pytest.mark.parametrize("scene", ["bunny", "sponza"])
def test_build_avx_flavour_fixtures(gpu, scene):
    from oracle import portpy, refpy
    v, label = scenes.load_scene(scene)
    o = refpy.RefBVH(v, mode=1, threaded=False) if refpy.available() else portpy.PortBVH(v, avx=True)
    e = api.BVH().BuildAVX(v)
    assert_same_tree(e, o.nodes, o.prim_idx, "BuildAVX " + label)
