#!/usr/bin/env python
"""Developer probe (NOT the bench): build one scene's SBVH on the GPU, make the three ray sets of SURVEY 8(d) (camera, shadow,
diffuse bounce) and time the device-resident traversal of one layout.  Small enough to sit under `ncu -k regex:k_trace`.

  python tools/trace_once.py <scene> <res> <bvh|cwbvh> [--sah] [--reps N] [--stats] [--sets primary,shadow,diffuse]
"""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import api, rays as R, scenes  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("res", type=int)
    ap.add_argument("layout", choices=["bvh", "cwbvh"])
    ap.add_argument("--sah", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--sets", default="primary,shadow,diffuse")
    a = ap.parse_args()
    v, label = scenes.load_scene(a.scene)
    e = api.BVH()
    (e.Build if a.sah else e.BuildHQ)(v)
    if a.layout == "cwbvh":
        api.check(api._lib.lib().tbvh_convert(e.h, api.LAYOUT_CWBVH))
        e.layout = api.LAYOUT_CWBVH
    i = e.info()
    print(f"{label}: {v.shape[0] // 3} tris, {'Build' if a.sah else 'BuildHQ'} {i.build_ms:.2f} ms, nodes {i.used_nodes}, depth {i.max_depth}, cwbvh nodes {i.used_blocks // 5}")
    lo, hi = scenes.scene_bounds(v)
    if a.scene == "sponza":
        eye, view = R.SPONZA_EYES[0], R.SPONZA_VIEWS[0]
    else:
        eye, view = R.bounds_camera(lo, hi, "inside" if a.scene == "bistro" else "outside")
    prim = R.primary_rays(eye, view, a.res, a.res, 16)
    n = prim.shape[0]
    dev = lambda r: torch.from_numpy(R.gpu_records(r).view(np.uint8).reshape(-1, 64)).cuda()
    hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
    dprim = dev(prim)
    e.Intersect(dprim, hits=hits)
    torch.cuda.synchronize()
    h = hits.cpu().numpy()
    traced = prim.copy()
    traced["t"], traced["u"], traced["v"], traced["prim"] = h[:, 0], h[:, 1], h[:, 2], h[:, 3].view(np.uint32)
    eps = float((hi - lo).max() * 5e-7)
    light = (lo + hi) * 0.5 + np.array([0, (hi - lo)[1] * 0.45, 0], np.float32) if a.scene != "sponza" else np.zeros(3, np.float32)
    sets = {"primary": (dprim, False)}
    want = a.sets.split(",")
    if "shadow" in want:
        sets["shadow"] = (dev(R.shadow_rays(traced, light, eps)), True)
    if "diffuse" in want:
        sets["diffuse"] = (dev(R.diffuse_rays(traced, v)), False)
    for name in want:
        d, anyhit = sets[name]
        fn = (lambda: e.IsOccluded(d, bits=bits)) if anyhit else (lambda: e.Intersect(d, hits=hits))
        extra = ""
        if a.stats:
            e.set_stats(True)
            fn()
            steps, tris = e.get_stats()[:2]
            e.set_stats(False)
            extra = f"  {steps / n:.2f} node visits/ray {tris / n:.2f} tri tests/ray"
        best, med = timeit(fn, a.reps)
        print(f"{name:8s} {'anyhit ' if anyhit else 'closest'} {n} rays: best {best:.3f} ms = {n / best / 1e3:.1f} Mrays/s (median {n / med / 1e3:.1f}){extra}")


if __name__ == "__main__":
    main()
