"""GPU parity: the CUDA BVH2 traversal (through the C-ABI) against the oracle - committed golden vectors, then the
oracle itself on seeded inputs, then fixtures at the reference's sizes.  Bar: bit-exact t,u,v,prim and occlusion bits
when the engine walks the oracle's own tree."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import golden_util as G
from tests import util

pytestmark = pytest.mark.gpu
ZERO = {"prim": 0, "t": 0, "u": 0, "v": 0}


def engine_from(nodes, prim_idx, verts, as_gpu_layout=None):
    if as_gpu_layout is not None:
        return api.BVH_GPU().upload(as_gpu_layout, prim_idx, verts)
    return api.BVH().upload(nodes, prim_idx, verts)


@pytest.mark.parametrize("path", G.golden_files(), ids=lambda p: p.split("/")[-1])
@pytest.mark.parametrize("layout", ["bvh", "bvh_gpu"])
def test_golden_vectors(gpu, path, layout):
    g = G.load(path)
    nodes = g["nodes"].view(np.uint8).view(api.NODE32).reshape(-1)
    gl = g["nodes_gpu"].view(np.uint8).view(api.NODE64).reshape(-1) if layout == "bvh_gpu" else None
    e = engine_from(nodes, g["prim_idx"], g["verts"], gl)
    for kind in ("primary", "diffuse"):
        r = G.rays_of(g, kind)
        e.Intersect(r)
        assert np.array_equal(G.hits_as_u32(r), g[kind + "_hit"]), f"{kind}: differs from BVH::Intersect"
    bits = e.IsOccluded(G.rays_of(g, "shadow"))
    assert np.array_equal(bits, g["shadow_bits"])


@pytest.mark.parametrize("ntris,seed,res", [(50000, 21, 128), (1500, 22, 64), (3, 23, 32)])
def test_seeded_scene_against_oracle(gpu, ntris, seed, res):
    v = scenes.procedural_scene(ntris, seed)
    o = util.oracle_bvh(v)
    e = api.BVH().upload(o.nodes, o.prim_idx, v)
    sets, bounds = util.ray_sets(v, res=res)
    want, got = sets["primary"].copy(), sets["primary"].copy()
    o.intersect(want), e.Intersect(got)
    assert util.compare_hits(got, want) == ZERO
    for name, rr in util.derived_sets(want, v, bounds).items():
        if name == "shadow":
            assert np.array_equal(e.IsOccluded(rr), o.occluded(rr))
        else:
            a, b = rr.copy(), rr.copy()
            o.intersect(a), e.Intersect(b)
            assert util.compare_hits(b, a) == ZERO


def test_device_path_equals_host_path_and_ragged_sizes(gpu):
    import torch
    v = scenes.procedural_scene(4000, 5)
    o = util.oracle_bvh(v)
    e = api.BVH().upload(o.nodes, o.prim_idx, v)
    sets, _ = util.ray_sets(v, res=64)
    for n in (1, 31, 33, 1000, sets["primary"].shape[0]):
        r = sets["primary"][:n].copy()
        want = r.copy()
        o.intersect(want)
        if n:
            e.Intersect(r)
            assert util.compare_hits(r, want) == ZERO
        # device path: packed 64-byte records, separate 16-byte hits
        d = torch.from_numpy(R.gpu_records(sets["primary"][:n]).view(np.uint8).reshape(-1, 64).copy()).cuda()
        hits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        e.Intersect(d, hits=hits)
        torch.cuda.synchronize()
        h = hits.cpu().numpy()
        assert np.array_equal(h[:, 0].view(np.uint32), want["t"].view(np.uint32))
        assert np.array_equal(h[:, 3].view(np.uint32), want["prim"])
        # in-place device hits and occlusion bits (ragged tail word)
        e.Intersect(d)
        bits = e.IsOccluded(d.clone())  # traced records: tmax = hit distance -> self-occlusion at t==tmax is accepted
        torch.cuda.synchronize()
        assert bits.shape[0] == (n + 31) // 32


def test_stats_counters(gpu):
    v = scenes.procedural_scene(4000, 5)
    o = util.oracle_bvh(v)
    e = api.BVH().upload(o.nodes, o.prim_idx, v)
    sets, _ = util.ray_sets(v, res=32)
    e.set_stats(True)
    e.Intersect(sets["primary"].copy())
    steps, tris = e.get_stats()[:2]
    assert steps > sets["primary"].shape[0] and tris > 0
    e.set_stats(False)


@pytest.mark.parametrize("scene", ["bunny", "sponza"])
def test_reference_fixtures_full_parity(gpu, scene):
    v, label = scenes.load_scene(scene)
    o = util.oracle_bvh(v)
    e = api.BVH().upload(o.nodes, o.prim_idx, v)
    lo, hi = scenes.scene_bounds(v)
    cams = [(R.SPONZA_EYES[i], R.SPONZA_VIEWS[i]) for i in range(3)] if scene == "sponza" else [R.bounds_camera(lo, hi, "outside")]
    for eye, view in cams:
        want = R.primary_rays(eye, view, 256, 256, 4)
        got = want.copy()
        o.intersect(want), e.Intersect(got)
        assert util.compare_hits(got, want) == ZERO, label
        d = util.derived_sets(want, v, (lo, hi))
        assert np.array_equal(e.IsOccluded(d["shadow"]), o.occluded(d["shadow"]))
        a, b = d["diffuse"].copy(), d["diffuse"].copy()
        o.intersect(a), e.Intersect(b)
        assert util.compare_hits(b, a) == ZERO, label


def test_traversal_cost_equals_the_sum_of_the_references_return_values(gpu):
    """BVH::Intersect returns (int32_t)( c_trav * nodes visited + c_int * triangles tested ) (tiny_bvh.h:3303); the speedtest sums it
    into rayCost.  The engine's counters (tbvh_get_stats) over a batch are that sum exactly (c_trav = c_int = 1)."""
    from oracle import refpy
    if not refpy.available():
        pytest.skip("needs oracle/_ref")
    v = scenes.procedural_scene(20000, 41)
    o = refpy.RefBVH(v, mode=0, threaded=False)
    e = api.BVH().upload(o.nodes, o.prim_idx, v)
    sets, bounds = util.ray_sets(v, res=128)
    rays = sets["primary"]
    want = o.intersect_cost(rays.copy())
    e.set_stats(True)
    e.Intersect(rays.copy())
    steps, tris = e.get_stats()[:2]
    e.set_stats(False)
    assert steps + tris == want, (steps, tris, want)
