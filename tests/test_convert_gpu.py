"""GPU parity of the layout conversions: BVH -> BVH_GPU on the device must equal BVH_GPU::ConvertFrom byte for byte."""
import numpy as np
import pytest

from oracle import refpy
from tinybvh_b200 import api, scenes
from tests import golden_util as G
from tests import util

pytestmark = pytest.mark.gpu


def diff_nodes(got, want, words):
    a, b = got.view(np.uint32).reshape(-1, words), np.ascontiguousarray(want).view(np.uint32).reshape(-1, words)
    assert a.shape == b.shape, f"node count {a.shape[0]} != {b.shape[0]}"
    bad = np.nonzero((a != b).any(1))[0]
    assert bad.size == 0, f"{bad.size} nodes differ, first {bad[:4]}: got {a[bad[0]]} want {b[bad[0]]}"


@pytest.mark.parametrize("path", G.golden_files(), ids=lambda p: p.split("/")[-1])
def test_bvh_gpu_conversion_matches_golden(gpu, path):
    g = G.load(path)
    e = api.BVH_GPU().Build(g["verts"])
    diff_nodes(e.download(), g["nodes_gpu"].view(np.uint8).view(api.NODE64).reshape(-1), 16)


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
@pytest.mark.parametrize("ntris,seed", [(5, 51), (999, 52), (120000, 53)])
def test_bvh_gpu_conversion_matches_reference(gpu, ntris, seed):
    v = scenes.procedural_scene(ntris, seed)
    ref = refpy.RefBVH(v, mode=0, threaded=False)
    want = refpy.RefBVHGPU(ref).nodes
    e = api.BVH_GPU().Build(v)
    diff_nodes(e.download(), want, 16)
    # the converted layout traverses like the source tree
    sets, _ = util.ray_sets(v, res=48)
    a, b = sets["primary"].copy(), sets["primary"].copy()
    ref.intersect(a), e.Intersect(b)
    assert util.compare_hits(b, a) == {"prim": 0, "t": 0, "u": 0, "v": 0}


@pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref")
def test_bvh_gpu_conversion_of_threaded_reference_tree(gpu):
    """An uploaded tree with the reference's threaded (non-DFS) node numbering converts to the same DFS layout."""
    import ctypes as C
    from tinybvh_b200 import _lib
    v = scenes.procedural_scene(90000, 54)
    ref = refpy.RefBVH(v, mode=0, threaded=True)
    want = refpy.RefBVHGPU(ref).nodes
    e = api.BVH_GPU()
    nodes, idx = np.ascontiguousarray(ref.nodes), np.ascontiguousarray(ref.prim_idx)
    api.check(_lib.lib().tbvh_upload_bvh(e.h, nodes.ctypes.data_as(C.c_void_p), nodes.shape[0], idx.ctypes.data_as(C.c_void_p), idx.shape[0],
                                         v.ctypes.data_as(C.c_void_p), 16, v.shape[0] // 3, api.HOST))
    api.check(_lib.lib().tbvh_convert(e.h, api.LAYOUT_BVH_GPU))
    diff_nodes(e.download(), want, 16)
