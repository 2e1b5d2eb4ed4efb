#!/usr/bin/env python
"""Developer probe: the one-process group API (tbvh_group_*) over every device of the box - Bistro CWBVH replicated peer-to-peer, one
host ray batch sharded by index; prints the replication time and the host-buffer throughput against one device."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import api, rays as R, scenes  # noqa: E402

v, label = scenes.load_scene("bistro")
e = api.BVH8_CWBVH().BuildHQ(v)
g = api.Group()
ms = g.replicate(e)
i = e.info()
print(f"{label}: {len(g)} devices, replicate {ms:.2f} ms for {(i.used_blocks // 5 * (80 + 160) + i.cwbvh_tri_count * 48 + i.used_nodes * 32 + i.idx_count * 52 + i.prim_count * 48) / 1e6:.0f} MB per device")
lo, hi = scenes.scene_bounds(v)
n = 1 << 25
rays = g.empty_rays(n, R.RAY_DTYPE)
R.primary_rays_into(rays, *R.bounds_camera(lo, hi, "inside"), 2048, 2048, 16)
one = api.pinned_empty(n, R.RAY_DTYPE, device=0)
one[:] = rays
bits = np.zeros((n + 31) // 32, np.uint32)
for name, fn in (("1 device  intersect", lambda: e.Intersect(one)), (f"{len(g)} devices intersect", lambda: g.Intersect(rays)),
                 ("1 device  occluded ", lambda: e.IsOccluded(one, bits=bits)), (f"{len(g)} devices occluded ", lambda: g.IsOccluded(rays, bits=bits))):
    fn()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    t = (time.perf_counter() - t0) / 3
    print(f"{name}: {t * 1e3:8.1f} ms  {n / t / 1e6:8.0f} Mrays/s (host buffers, {n} camera rays)", flush=True)
assert np.array_equal(one["t"].view(np.uint32), rays["t"].view(np.uint32)) and np.array_equal(one["prim"], rays["prim"]), "group hits differ from one device"
print("group hits identical to one device")
