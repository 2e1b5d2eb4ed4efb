"""Every tuning variant (tbvh_set_option) must give the oracle's results: the traversal-kernel variants change which
lane / code path runs a ray, never its arithmetic or order; the host-path modes change how bytes cross PCIe."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import util

pytestmark = pytest.mark.gpu
ZERO = {"prim": 0, "t": 0, "u": 0, "v": 0}


@pytest.fixture(scope="module")
def world():
    v = scenes.procedural_scene(40000, 71)
    o = util.oracle_bvh(v)
    sets, bounds = util.ray_sets(v, res=128)
    prim = sets["primary"].copy()
    o.intersect(prim)
    d = util.derived_sets(prim, v, bounds)
    want = {"primary": prim, "diffuse": d["diffuse"].copy(), "shadow_bits": o.occluded(d["shadow"])}
    o.intersect(want["diffuse"])
    return v, sets["primary"], d, want


@pytest.mark.parametrize("variant", [0, 3, 4])
def test_trace_variants_are_bit_exact(gpu, world, variant):
    v, primary, d, want = world
    import torch
    api.set_option("trace_variant", variant)
    try:
        e = api.BVH().Build(v)
        for name, rays in (("primary", primary), ("diffuse", d["diffuse"])):
            got = rays.copy()
            e.Intersect(got)
            assert util.compare_hits(got, want[name]) == ZERO, f"variant {variant} {name}"
            # device path with a ragged count (persistent kernel tail handling)
            m = rays.shape[0] - 37
            dev = torch.from_numpy(R.gpu_records(rays[:m]).view(np.uint8).reshape(-1, 64).copy()).cuda()
            hits = torch.zeros((m, 4), dtype=torch.float32, device="cuda")
            e.Intersect(dev, hits=hits)
            torch.cuda.synchronize()
            h = hits.cpu().numpy()
            assert np.array_equal(h[:, 0].view(np.uint32), want[name]["t"][:m].view(np.uint32))
            assert np.array_equal(h[:, 3].view(np.uint32), want[name]["prim"][:m])
        assert np.array_equal(e.IsOccluded(d["shadow"]), want["shadow_bits"]), f"variant {variant} occlusion"
        m = d["shadow"].shape[0] - 37
        assert np.array_equal(e.IsOccluded(d["shadow"][:m].copy()), util.oracle_bvh(v).occluded(d["shadow"][:m].copy()))
    finally:
        api.set_option("trace_variant", 3)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_host_path_modes_return_the_same_hits(gpu, world, mode):
    v, primary, d, want = world
    api.set_option("d2h_mode", mode)
    try:
        e = api.BVH().Build(v)
        n = primary.shape[0]
        pinned = api.pinned_empty(n, R.RAY_DTYPE)
        pinned[:] = primary
        e.Intersect(pinned)
        assert util.compare_hits(pinned, want["primary"]) == ZERO, f"d2h_mode {mode} (pinned)"
        pageable = primary.copy()
        e.Intersect(pageable)
        assert util.compare_hits(pageable, want["primary"]) == ZERO, f"d2h_mode {mode} (pageable)"
        hits = e.IntersectPacked(primary)   # packed return path: rays untouched, 16-byte hits
        assert np.array_equal(hits["t"].view(np.uint32), want["primary"]["t"].view(np.uint32)) and np.array_equal(hits["prim"], want["primary"]["prim"])
        api.pinned_free(pinned)
    finally:
        api.set_option("d2h_mode", 0)


@pytest.mark.parametrize("small_t", [8, 64, 256])
def test_builder_switch_point_does_not_change_the_tree(gpu, small_t):
    v = scenes.procedural_scene(30000, 72)
    o = util.oracle_bvh(v)
    api.set_option("small_t", small_t)
    try:
        nodes, idx = api.BVH().Build(v).download()
        assert np.array_equal(nodes.view(np.uint32), np.ascontiguousarray(o.nodes).view(np.uint32)) and np.array_equal(idx, o.prim_idx)
    finally:
        api.set_option("small_t", 128)


def test_packed_upload_host_path(gpu):
    """host_path 2: host threads pack O, D, mask, t (u, v, prim) into 48 / 32 bytes per ray after checking rD == safercp( D ),
    the device rebuilds rD.  Same bits as the default path; rays with a hand-set rD fall back to the 64-byte copy; a miss leaves
    u, v, prim as they were."""
    v = scenes.procedural_scene(60000, 77)
    o = util.oracle_bvh(v)
    lo, hi = scenes.scene_bounds(v)
    rays = R.primary_rays(*R.bounds_camera(lo, hi, "outside"), 512, 512, 4)      # 1,048,576 rays: above the 65,536-ray threshold
    rays["u"], rays["v"], rays["prim"] = 0.25, 0.5, 77                            # what a miss must leave behind
    want = rays.copy()
    o.intersect(want)
    sh = util.derived_sets(want, v, (lo, hi))["shadow"]
    want_bits = o.occluded(sh)
    e = api.BVH().Build(v)
    api.set_option("host_path", 2)
    try:
        got = rays.copy()
        e.Intersect(got)
        assert util.compare_hits(got, want) == ZERO
        miss = want["t"] >= 1e30
        assert miss.any() and np.all(got["prim"][miss] == 77) and np.all(got["u"][miss] == np.float32(0.25))
        assert np.array_equal(e.IsOccluded(sh), want_bits)
        hits = e.IntersectPacked(rays)
        assert np.array_equal(hits["t"].view(np.uint32), want["t"].view(np.uint32)) and np.array_equal(hits["prim"], want["prim"])
        # a chunk holding a ray whose rD is not safercp( D ) must take the 64-byte path and honour the stored rD
        odd = rays.copy()
        odd["rD"][123456] = odd["rD"][123456] * np.float32(1.5)
        want_odd = odd.copy()
        o.intersect(want_odd)
        e.Intersect(odd)
        assert util.compare_hits(odd, want_odd) == ZERO
    finally:
        api.set_option("host_path", 0)
