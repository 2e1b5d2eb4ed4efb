#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref (needs /root/reference at build
time; run in the build container).  Each file holds a small triangle soup, ray sets, and what the reference's
BVH::Build + BVH::Intersect / IsOccluded produce for them: the node array, primIdx, per-ray (t,u,v,prim) and
occlusion bits.  The restatement (oracle/tbvh_oracle.c) and the CUDA engine are both tested against these."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import rays as R, scenes  # noqa: E402
from oracle import refpy  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def ray_core(r):
    return {"O": r["O"].copy(), "D": r["D"].copy(), "rD": r["rD"].copy(), "tmax": r["t"].copy()}


def make(name, verts, res):
    ref = refpy.RefBVH(verts, mode=0, threaded=False)
    lo, hi = scenes.scene_bounds(verts)
    prim = np.concatenate([R.primary_rays(*R.bounds_camera(lo, hi, k), res, res, 4) for k in ("outside", "inside")])
    d = {"verts": verts, "nodes": ref.nodes.copy().view(np.uint32).reshape(-1, 8), "prim_idx": ref.prim_idx.copy()}
    for k, v in ray_core(prim).items():
        d["primary_" + k] = v
    ref.intersect(prim, threads=1)
    d["primary_hit"] = np.stack([prim["t"].view(np.uint32), prim["u"].view(np.uint32), prim["v"].view(np.uint32), prim["prim"]], 1)
    eps = float((hi - lo).max() * 5e-7)
    light = (lo + hi) * 0.5 + np.array([0, (hi - lo)[1] * 0.45, 0], np.float32)
    sh = R.shadow_rays(prim, light, eps)
    for k, v in ray_core(sh).items():
        d["shadow_" + k] = v
    d["shadow_bits"] = ref.occluded(sh, threads=1)
    df = R.diffuse_rays(prim, verts)
    for k, v in ray_core(df).items():
        d["diffuse_" + k] = v
    ref.intersect(df, threads=1)
    d["diffuse_hit"] = np.stack([df["t"].view(np.uint32), df["u"].view(np.uint32), df["v"].view(np.uint32), df["prim"]], 1)
    g = refpy.RefBVHGPU(ref)
    d["nodes_gpu"] = g.nodes.copy().view(np.uint32).reshape(-1, 16)
    # BVH::BuildHQ (SBVH), single-threaded numbering, after its closing Compact()
    hq = refpy.RefBVH(verts, mode=2, threaded=False)
    d["hq_nodes"] = hq.nodes.copy().view(np.uint32).reshape(-1, 8)
    d["hq_prim_idx"] = hq.prim_idx[: int(hq.nodes["triCount"].sum())].copy()
    d["hq_idx_count"] = np.array([hq.idx_count], np.uint32)
    # BVH8_CWBVH: the conversion chain (SplitLeafs(3), MBVH<8>, CWBVH encode) over the BVH::Build tree, and the reference's CPU walk
    if verts.shape[0] > 3:   # "converting a single-node bvh" is a fatal error in the reference (:5889)
        cw = refpy.RefCWBVH(verts, mode=2)
        d["cwbvh_nodes"] = cw.nodes.copy().view(np.uint32)
        d["cwbvh_tris"] = cw.tris[: verts.shape[0]].copy().view(np.uint32)     # 3 records per triangle reference
        cp = R.primary_rays(*R.bounds_camera(lo, hi, "outside"), res, res, 4)
        cw.intersect(cp, threads=1)
        d["cwbvh_primary_hit"] = np.stack([cp["t"].view(np.uint32), cp["u"].view(np.uint32), cp["v"].view(np.uint32), cp["prim"]], 1)
    # BVH::Refit (:3055) after every vertex moved a little (same topology)
    rng = np.random.default_rng(97)
    w = verts.copy()
    w[:, :3] += (rng.random((verts.shape[0], 3), np.float32) - 0.5) * np.float32(0.02 * float((hi - lo).max()))
    rf = refpy.RefBVH(verts, mode=0, threaded=False)
    rf.refit(w)
    d["refit_verts"], d["refit_nodes"] = w, rf.nodes.copy().view(np.uint32).reshape(-1, 8)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "tris", verts.shape[0] // 3, "nodes", ref.used_nodes, "rays", prim.shape[0],
          "hits", int((prim["t"] < 1e30).sum()), "occluded", int(np.unpackbits(d["shadow_bits"].view(np.uint8)).sum()))


def make_tlas(name):
    """TLAS / BLAS golden vectors: BLASInstance::Update, BVH::Build( BLASInstance*, .. ), IntersectTLAS / IsOccludedTLAS of the reference."""
    from tests.util import random_transforms
    v = [scenes.procedural_scene(600, 31), scenes.procedural_scene(200, 32)]
    raw = refpy.make_instances(random_transforms(24, 33), [i % 2 for i in range(24)], masks=[0x3 if i % 5 else 0x2 for i in range(24)])
    inst = raw.copy()
    tl = refpy.RefTLAS(inst, [refpy.RefBVH(x, mode=0, threaded=False) for x in v])    # Update()s inst
    tb = tl.bvh()
    rng = np.random.default_rng(34)
    D = rng.normal(size=(6000, 3)).astype(np.float32) * 0.35 + np.array([0, 0, 1], np.float32)
    O = np.tile(np.array([[0, 0, -120]], np.float32), (D.shape[0], 1))
    d = {"verts0": v[0], "verts1": v[1], "instances_raw": raw.view(np.uint32).reshape(-1, 48), "instances": inst.view(np.uint32).reshape(-1, 48),
         "tlas_nodes": tb.nodes.copy().view(np.uint32).reshape(-1, 8), "tlas_prim_idx": tb.prim_idx.copy()}
    for mask in (1, 2):
        r = R.make_rays(O, D)
        r["mask"] = mask
        for k, x in ray_core(r).items():
            d[f"rays_{k}"] = x
        tl.intersect(r, threads=1)
        d[f"hit_mask{mask}"] = r.view(np.uint32).reshape(-1, 32)[:, 11:16].copy()      # inst, t, u, v, prim
        sh = R.make_rays(O, D, tmax=150.0)
        sh["mask"] = mask
        d[f"occluded_mask{mask}"] = tl.occluded(sh, threads=1)
    os.makedirs(os.path.join(OUT, "tlas"), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "tlas", name + ".npz"), **d)
    print(name, "instances", inst.shape[0], "tlas nodes", tb.used_nodes, "hits", int((d["hit_mask2"][:, 1].view(np.float32) < 1e30).sum()))


if __name__ == "__main__":
    make_tlas("tlas_24")
    make("atrium_3k", scenes.procedural_scene(3000, seed=11), 32)
    # degenerate inputs: coincident triangles (exact t ties), a flat (zero-extent axis) soup, a single triangle
    rng = np.random.default_rng(5)
    base = scenes.procedural_scene(400, seed=3)
    dup = np.concatenate([base, base[: 150 * 3], base[: 60 * 3]])
    make("coincident_610", dup, 24)
    flat = scenes.procedural_scene(500, seed=9)
    flat[:, 1] = 2.0
    make("flat_500", flat, 24)
    make("single_tri", np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0]], np.float32), 16)
