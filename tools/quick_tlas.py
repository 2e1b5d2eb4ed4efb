#!/usr/bin/env python
"""Developer probe: two-level traversal throughput.  A grid of instances of one BLAS (bunny), rotated and scaled, device-resident
camera rays; the reference's IntersectTLAS on all host threads beside it on a sample of the same rays."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import api, rays as R, scenes  # noqa: E402
from tools.quick_perf import timeit  # noqa: E402
from oracle import refpy  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    v, label = scenes.load_scene("bunny")
    lo, hi = scenes.scene_bounds(v)
    ext = float((hi - lo).max())
    rng = np.random.default_rng(7)
    T = np.zeros((side * side, 16), np.float32)
    for i in range(side * side):
        a = rng.random() * 6.28
        s = 0.6 + 0.6 * rng.random()
        m = np.eye(4)
        m[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) * s
        m[:3, 3] = [(i % side - side / 2) * ext * 1.3, 0, (i // side - side / 2) * ext * 1.3]
        T[i] = m.astype(np.float32).reshape(-1)
    inst = refpy.make_instances(T, np.zeros(side * side, np.uint32))
    rb = refpy.RefBVH(v, mode=0)
    ref = refpy.RefTLAS(inst, [rb])   # Update()s the instances
    blas = api.BVH().Build(v)
    t0 = time.time()
    tl = api.TLAS().Build(inst, [blas])
    print(f"{side * side} instances of {label} ({v.shape[0] // 3} tris each): TLAS build {tl.info().build_ms:.3f} ms device, {tl.info().used_nodes} nodes ({(time.time() - t0) * 1e3:.1f} ms wall)")
    span = side * ext * 1.3
    eye = np.array([0, span * 0.35, -span * 0.75], np.float32)
    target = np.array([0, 0, 0], np.float32)
    prim = R.primary_rays(eye, target - eye, res, res, 4)
    n = prim.shape[0]
    d = torch.from_numpy(prim.view(np.uint8).reshape(-1, 128)).cuda()
    fresh = d.clone()
    def run():
        d.copy_(fresh)
        tl.Intersect(d)
    tcopy, _ = timeit(lambda: d.copy_(fresh))
    best, med = timeit(run)
    got = d.cpu().numpy().view(R.RAY_DTYPE).reshape(-1)
    print(f"IntersectTLAS on the GPU: {best - tcopy:.3f} ms for {n} rays = {n / (best - tcopy) / 1e3:.1f} Mrays/s ({(got['t'] < 1e30).mean() * 100:.1f} % hit)")
    sh = prim.copy(); sh["t"] = span
    dsh = torch.from_numpy(sh.view(np.uint8).reshape(-1, 128)).cuda()
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
    best, _ = timeit(lambda: tl.IsOccluded(dsh, bits=bits))
    print(f"IsOccludedTLAS on the GPU: {best:.3f} ms = {n / best / 1e3:.1f} Mrays/s")
    # the same instances over the BVH8_CWBVH layout of the BLAS (the reference's GPU arrangement, traverse_tlas.cl)
    wide = api.BVH8_CWBVH().Build(v)
    tw = api.TLAS().Build(inst.copy(), [wide], blas_layout=api.LAYOUT_CWBVH)
    dw = fresh.clone()
    def run_wide():
        dw.copy_(fresh)
        tw.Intersect(dw)
    best, _ = timeit(run_wide)
    gw = dw.cpu().numpy().view(R.RAY_DTYPE).reshape(-1)
    ww = lambda r: r.view(np.uint32).reshape(-1, 32)[:, 11:16]
    print(f"TLAS over the CWBVH BLAS: {best - tcopy:.3f} ms = {n / (best - tcopy) / 1e3:.1f} Mrays/s; {int((ww(gw) != ww(got)).any(axis=1).sum())} of {n} rays differ from the BVH-BLAS walk (ties)")
    best, _ = timeit(lambda: tw.IsOccluded(dsh, bits=bits))
    print(f"... any-hit: {best:.3f} ms = {n / best / 1e3:.1f} Mrays/s")
    smp = prim[: 1 << 20].copy()
    t0 = time.time(); ref.intersect(smp, threads=0); dt = time.time() - t0
    print(f"reference IntersectTLAS, {refpy.hardware_threads()} host threads: {smp.shape[0] / dt / 1e6:.1f} Mrays/s")
    w = lambda r: r.view(np.uint32).reshape(-1, 32)[:, 11:16]
    print("sample identical to the reference:", bool(np.array_equal(w(got[: 1 << 20]), w(smp))))


if __name__ == "__main__":
    main()
