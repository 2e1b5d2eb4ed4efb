#!/usr/bin/env python
"""Developer probe: does ordering incoherent rays (origin cell Morton code + direction octant) before traversal pay?
Times k_trace_bvh2 on diffuse-bounce rays in generation order and in sorted order (the sort itself is timed apart)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import api, rays as R, scenes  # noqa: E402
from tools.quick_perf import timeit  # noqa: E402


def part1by2(x):
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "bistro"
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    v, label = scenes.load_scene(scene)
    e = api.BVH().Build(v)
    lo, hi = scenes.scene_bounds(v)
    eye, view = (R.SPONZA_EYES[0], R.SPONZA_VIEWS[0]) if scene == "sponza" else R.bounds_camera(lo, hi, "inside" if scene == "bistro" else "outside")
    prim = R.primary_rays(eye, view, res, res, 4)
    n = prim.shape[0]
    d = torch.from_numpy(R.gpu_records(prim).view(np.uint8).reshape(-1, 64)).cuda()
    hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    e.Intersect(d, hits=hits)
    h = hits.cpu().numpy()
    traced = prim.copy()
    traced["t"], traced["u"], traced["v"], traced["prim"] = h[:, 0], h[:, 1], h[:, 2], h[:, 3].view(np.uint32)
    df = R.diffuse_rays(traced, v)
    ddf = torch.from_numpy(R.gpu_records(df).view(np.uint8).reshape(-1, 64)).cuda()
    best, _ = timeit(lambda: e.Intersect(ddf, hits=hits))
    ref_hits = hits.clone()
    print(f"{label}: {n} diffuse rays, generation order: {best:.3f} ms = {n / best / 1e3:.1f} Mrays/s")
    f = ddf.view(torch.float32).reshape(-1, 16)
    O, D = f[:, 0:3], f[:, 4:7]
    tlo, thi = torch.tensor(lo, device="cuda"), torch.tensor(hi, device="cuda")
    for bits in (5, 7, 10):
        def key():
            q = ((O - tlo) / (thi - tlo) * (1 << bits)).clamp(0, (1 << bits) - 1).to(torch.int64)
            m = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)
            octant = (D[:, 0] < 0).to(torch.int64) | ((D[:, 1] < 0).to(torch.int64) << 1) | ((D[:, 2] < 0).to(torch.int64) << 2)
            return (octant << 30) | m, (m << 3) | octant
        for name, k in zip(("octant-major", "cell-major"), key()):
            ts, _ = timeit(lambda: torch.sort(k))
            order = torch.sort(k).indices
            tg, _ = timeit(lambda: ddf[order])
            srt = ddf[order].contiguous()
            best2, _ = timeit(lambda: e.Intersect(srt, hits=hits))
            ok = torch.equal(hits.view(torch.int32), ref_hits[order].view(torch.int32))
            print(f"  {bits:2d} bits/axis {name:12s}: trace {best2:.3f} ms = {n / best2 / 1e3:.1f} Mrays/s ({best / best2:.2f}x), torch.sort {ts:.3f} ms, gather {tg:.3f} ms, hits identical {ok}")


if __name__ == "__main__":
    main()
