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
    # (a) same data, same order: identical primitive; t,u,v bit-identical where the oracle pairing is shared
    a = util.compare_hits(got, want_cw)
    assert a["prim"] == 0, f"{label}: prim differs from BVH8_CWBVH::Intersect on {a['prim']} rays"
    rel = np.abs(got["t"] - want_cw["t"]) / np.maximum(np.abs(want_cw["t"]), 1e-30)
    assert (rel[want_cw["t"] < 1e30] <= 1e-6).all()
    # (b) oracle
    hit_o, hit_g = want_o["t"] < 1e30, got["t"] < 1e30
    assert np.array_equal(hit_o, hit_g), f"{label}: hit/miss flips vs oracle: {(hit_o != hit_g).sum()}"
    rel = np.abs(got["t"] - want_o["t"]) / np.maximum(np.abs(want_o["t"]), 1e-30)
    assert (rel[hit_o] <= 1e-4).all(), f"{label}: t off by more than 1e-4 relative"
    cls = util.classify_mismatches(got, want_o, verts)
    assert cls["real"] == 0, f"{label}: {cls}"
    same = got["prim"] == want_o["prim"]
    assert np.array_equal(got["t"][same & hit_o].view(np.uint32), want_o["t"][same & hit_o].view(np.uint32)), f"{label}: same prim, different t bits"
    return cls


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
    # occlusion: identical bits to the oracle except rays whose only blockers are grazing ties; demand equality here
    assert np.array_equal(e.IsOccluded(d["shadow"]), o.occluded(d["shadow"]))


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
    got, want = e.IsOccluded(d["shadow"]), o.occluded(d["shadow"])
    diff = int(np.unpackbits((got ^ want).view(np.uint8)).sum())
    assert diff == 0, f"{label}: {diff} occlusion bits differ from BVH::IsOccluded"
