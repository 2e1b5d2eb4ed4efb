"""Shared helpers for the parity tests: oracle selection, ray sets, comparators."""
import numpy as np

from tinybvh_b200 import rays as R, scenes
from oracle import portpy, refpy


def oracle_bvh(verts):
    """The parity oracle for a triangle soup: the compiled reference (BVH::Build + Intersect/IsOccluded) when
    oracle/_ref exists, else the pinned plain-C restatement.  Both expose .nodes/.prim_idx/.intersect/.occluded."""
    if refpy.available():
        return refpy.RefBVH(verts, mode=0, threaded=False)
    return portpy.PortBVH(verts)


def small_scene(ntris=6000, seed=7):
    return scenes.procedural_scene(ntris, seed)


def ray_sets(verts, res=96, seed=0x123456):
    """primary (2 cameras) + shadow + diffuse rays over a scene; dict name -> RAY_DTYPE array (untraced)."""
    lo, hi = scenes.scene_bounds(verts)
    out = {}
    prim = []
    for kind in ("outside", "inside"):
        eye, view = R.bounds_camera(lo, hi, kind)
        prim.append(R.primary_rays(eye, view, res, res, 4))
    out["primary"] = np.concatenate(prim)
    return out, (lo, hi)


def derived_sets(traced_primary, verts, bounds):
    lo, hi = bounds
    eps = float((hi - lo).max() * 5e-7)
    light = (lo + hi) * 0.5 + np.array([0, (hi - lo)[1] * 0.45, 0], np.float32)
    return {"shadow": R.shadow_rays(traced_primary, light, eps), "diffuse": R.diffuse_rays(traced_primary, verts)}


def bits_u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


def compare_hits(got, want):
    """-> dict of mismatch counts between two traced ray arrays (bit-exact fields)."""
    return {
        "prim": int((got["prim"] != want["prim"]).sum()),
        "t": int((bits_u32(got["t"]) != bits_u32(want["t"])).sum()),
        "u": int((bits_u32(got["u"]) != bits_u32(want["u"])).sum()),
        "v": int((bits_u32(got["v"]) != bits_u32(want["v"])).sum()),
    }


def classify_mismatches(got, want, verts):
    """Tie audit: for rays whose prim differs from the oracle's, re-evaluate the engine's prim with the
    oracle's Moeller-Trumbore arithmetic.  exact-tie = bit-identical t; otherwise 'real'."""
    bad = np.nonzero(got["prim"] != want["prim"])[0]
    ties = real = 0
    v = verts.reshape(-1, 3, 4)
    for i in bad:
        p = int(got["prim"][i])
        ok, t, u, vv = portpy.tri_test(want["O"][i], want["D"][i], v[p, 0, :3], v[p, 1, :3], v[p, 2, :3], 1e30)
        if ok and np.float32(t).view(np.uint32) == want["t"][i].view(np.uint32):
            ties += 1
        else:
            real += 1
    return {"mismatch": int(bad.size), "tie_equivalent": ties, "real": real}


def random_transforms(count, seed, spread=60.0):
    """Row-major 4x4 instance transforms: rotation x (non-uniform) scale x translation, as tiny_bvh's bvhmat4 stores them."""
    rng = np.random.default_rng(seed)
    out = np.zeros((count, 16), np.float32)
    for i in range(count):
        a, b, c = rng.random(3) * 6.28
        rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        m = np.eye(4)
        m[:3, :3] = (rx @ ry @ rz) @ np.diag(0.4 + rng.random(3) * 1.2)
        m[:3, 3] = (rng.random(3) - 0.5) * spread
        out[i] = m.astype(np.float32).reshape(-1)
    return out
