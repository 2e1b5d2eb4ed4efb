"""Several devices in one process (tbvh_group_*): a BVH replicated over the group, a host ray batch sharded by index.
On a one-GPU box the group is built over device 0 twice - two contexts, two replicas, the same code path as two GPUs."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import util

pytestmark = pytest.mark.gpu
ZERO = {"prim": 0, "t": 0, "u": 0, "v": 0}


def devices():
    n = api.device_count()
    return list(range(n)) if n > 1 else [0, 0]


def test_shard_range_matches_the_python_partition():
    from tinybvh_b200 import multi
    for n in (0, 1, 31, 32, 33, 1000, 1 << 20, (1 << 20) + 17):
        for parts in (1, 2, 3, 8):
            got = [api.shard_range(n, p, parts) for p in range(parts)]
            assert got == [multi.shard_range(n, p, parts) for p in range(parts)]
            assert sum(c for _, c in got) == n and all(a % 32 == 0 for a, c in got if c)


@pytest.mark.parametrize("layout", ["bvh", "cwbvh"])
def test_group_traversal_equals_one_device(gpu, layout):
    v = scenes.procedural_scene(25000, 91)
    lo, hi = scenes.scene_bounds(v)
    e = api.BVH().BuildHQ(v)
    if layout == "cwbvh":
        api.check(api._lib.lib().tbvh_convert(e.h, api.LAYOUT_CWBVH))
        e.layout = api.LAYOUT_CWBVH
    g = api.Group(devices())
    ms = g.replicate(e)
    assert ms >= 0 and len(g) == len(devices())
    n = 300_017                      # not a multiple of 32 x parts
    src = R.primary_rays(*R.bounds_camera(lo, hi, "inside"), 160, 160, 16)[:n]
    rays = g.empty_rays(n, R.RAY_DTYPE)
    rays[:] = src
    want = src.copy()
    e.Intersect(want)
    g.Intersect(rays)
    assert util.compare_hits(rays, want) == ZERO
    sh = util.derived_sets(want, v, (lo, hi))["shadow"]
    assert np.array_equal(g.IsOccluded(sh), e.IsOccluded(sh))
    # the source changes: replicate again
    e2 = api.BVH().Build(v)
    g.replicate(e2)
    got2, want2 = src.copy(), src.copy()
    e2.Intersect(want2), g.Intersect(got2)
    assert util.compare_hits(got2, want2) == ZERO
    g.close()


def test_group_needs_a_replica(gpu):
    g = api.Group([0])
    with pytest.raises(api.TbvhError):
        g.Intersect(R.make_rays(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32)))
    g.close()
