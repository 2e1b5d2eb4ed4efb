"""GPU parity of BVH::Refit (tiny_bvh.h:3055): after the vertices moved, the refitted node array is the reference's byte for
byte and traversal of the refitted tree is bit-exact."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import util
from tests.test_oracle_pin import moved

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,method", [("synthetic:1", "Build"), ("synthetic:5000", "Build"), ("synthetic:70000", "BuildAVX"), ("sponza", "Build")])
def test_refit_matches_oracle(gpu, scene, method):
    from oracle import portpy
    v, label = scenes.load_scene(scene)
    o = portpy.PortBVH(v, avx=method == "BuildAVX")
    e = getattr(api.BVH(), method)(v)
    for frame in (1, 2):  # refit twice: the second starts from refitted boxes
        w = moved(v, 60 + frame)
        o.refit(w), e.Refit(w)
        nodes, idx = e.download()
        assert np.array_equal(nodes.view(np.uint32), o.nodes.view(np.uint32)), f"{label}: refitted nodes differ (frame {frame})"
        assert np.array_equal(idx, o.prim_idx)
    assert e.info().build_ms > 0
    lo, hi = scenes.scene_bounds(w)
    eye, view = (R.SPONZA_EYES[0], R.SPONZA_VIEWS[0]) if scene == "sponza" else R.bounds_camera(lo, hi, "outside")
    want = R.primary_rays(eye, view, 96, 96, 4)
    got = want.copy()
    o.intersect(want), e.Intersect(got)
    assert util.compare_hits(got, want) == {"prim": 0, "t": 0, "u": 0, "v": 0}
    print(f"{label}: refit {e.info().build_ms:.3f} ms")


def test_refit_refuses_sbvh_and_drops_derived_layouts(gpu):
    v = scenes.procedural_scene(3000, 71)
    with pytest.raises(api.TbvhError):
        api.BVH().BuildHQ(v).Refit(v)   # "BVH::Refit( .. ), refitting an SBVH." (:3057)
    g = api.BVH_GPU().Build(v)
    assert g.info().layouts & (1 << api.LAYOUT_BVH_GPU)
    api.BVH.Refit(g, moved(v, 72))
    assert g.info().layouts == 1 << api.LAYOUT_BVH
