"""GPU parity of the two-level path: BVH::Build( BLASInstance*, .. ) (tiny_bvh.h:2221) + IntersectTLAS (:3306) / IsOccludedTLAS
(:3455).  The TLAS node array is the reference's byte for byte; hits (inst, t, u, v, prim) and occlusion bits are identical."""
import numpy as np
import pytest

from tinybvh_b200 import api, rays as R, scenes
from tests import util
from tests.test_oracle_pin import tlas_case

pytestmark = pytest.mark.gpu


def words(r):
    return r.view(np.uint32).reshape(-1, 32)[:, 11:16]   # hit.inst, t, u, v, prim


@pytest.mark.parametrize("n_inst,builder", [(40, "Build"), (1, "Build"), (300, "BuildAVX"), (40, "BuildHQ")])
def test_tlas_matches_reference(gpu, n_inst, builder):
    from oracle import refpy
    if not refpy.available():
        pytest.skip("needs oracle/_ref")
    v, inst, O, D = tlas_case(91, n_inst)
    mode = {"Build": 0, "BuildAVX": 1, "BuildHQ": 2}[builder]
    inst_ref = inst.copy()
    ref = refpy.RefTLAS(inst_ref, [refpy.RefBVH(x, mode=mode, threaded=False) for x in v])   # Update()s inst_ref in place
    blas = [getattr(api.BVH(), builder)(x) for x in v]
    t = api.TLAS().Build(inst, blas)                                                         # the engine Update()s inst in place
    assert inst.tobytes() == inst_ref.tobytes(), "BLASInstance::Update differs"
    nodes, idx = t.download()
    rb = ref.bvh()
    assert np.array_equal(nodes.view(np.uint32), rb.nodes.view(np.uint32)) and np.array_equal(idx, rb.prim_idx), "TLAS tree differs"
    for mask in (0x1, 0x2):
        rays = R.make_rays(O, D)
        rays["mask"] = mask
        want, got = rays.copy(), rays.copy()
        ref.intersect(want), t.Intersect(got)
        assert np.array_equal(words(got), words(want)), f"closest hits differ (ray mask {mask:#x})"
        sh = R.make_rays(O, D, tmax=150.0)
        sh["mask"] = mask
        assert np.array_equal(t.IsOccluded(sh), ref.occluded(sh))
    hit = want["t"] < 1e30
    assert hit.sum() > 1000


def test_tlas_device_rays_and_errors(gpu):
    import torch
    from oracle import refpy
    if not refpy.available():
        pytest.skip("needs oracle/_ref")
    v, inst, O, D = tlas_case(93, 24)
    ref = refpy.RefTLAS(inst, [refpy.RefBVH(x, mode=0, threaded=False) for x in v])
    blas = [api.BVH().Build(x) for x in v]
    t = api.TLAS().Build(inst, blas, update=False)   # records already updated by the reference: the "blasses == 0" contract
    rays = R.make_rays(O, D)
    want = rays.copy()
    ref.intersect(want)
    d = torch.from_numpy(rays.view(np.uint8).reshape(-1, 128)).cuda()
    t.Intersect(d)
    got = d.cpu().numpy().view(R.RAY_DTYPE).reshape(-1)
    assert np.array_equal(words(got), words(want))
    with pytest.raises(api.TbvhError):
        t.IntersectPacked(rays)                      # TLAS hits carry the instance: in place only
    with pytest.raises(api.TbvhError):
        api.TLAS().Build(inst, [blas[0]])            # an instance names BLAS 1
    with pytest.raises(api.TbvhError):
        api.TLAS().Build(inst, [blas[0], t])         # a TLAS is not a BLAS


def test_tlas_instance_bits_in_prim(gpu):
    """A host program compiled with INST_IDX_BITS 10 (the speedtest's setting): the instance rides in the top bits of hit.prim."""
    from oracle import refpy
    if not refpy.available():
        pytest.skip("needs oracle/_ref")
    v, inst, O, D = tlas_case(95, 30)
    ref = refpy.RefTLAS(inst, [refpy.RefBVH(x, mode=0, threaded=False) for x in v])
    t = api.TLAS().Build(inst, [api.BVH().Build(x) for x in v])
    rays = R.make_rays(O, D)
    want, got = rays.copy(), rays.copy()
    ref.intersect(want)                                   # reference built with INST_IDX_BITS 32: inst in its own field
    api.set_option("inst_idx_bits", 10)
    try:
        t.Intersect(got)
    finally:
        api.set_option("inst_idx_bits", 32)
    hit = want["t"] < 1e30
    w, g = words(want), words(got)
    assert np.array_equal(g[:, 1:4], w[:, 1:4])           # t, u, v
    assert np.array_equal(g[hit, 4], w[hit, 4] + (w[hit, 0] << 22)) and np.array_equal(g[~hit, 4], w[~hit, 4])
    assert np.array_equal(g[:, 0], words(rays)[:, 0])     # byte 44 untouched
