#!/usr/bin/env python
"""Developer probe: which host-buffer strategy moves 64 B of every 128-byte host Ray fastest over PCIe?  A one-triangle
scene makes traversal free, so tbvh_intersect's time is the transfer time.  Run once per mode (env is read at context
creation):  TBVH_HOST_PATH=copy2d|zerocopy  TBVH_H2D_SPLIT=1..4  python tools/pcie_probe.py"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tinybvh_b200 import api, rays as R  # noqa: E402

n = 1 << 24
v = np.array([[-10, -10, 5, 0], [10, -10, 5, 0], [0, 10, 5, 0]], np.float32)  # every ray hits at t = 5
e = api.BVH().Build(v)
h = api.pinned_empty(n, R.RAY_DTYPE)
h[:] = R.make_rays(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32))[0]
bits = np.zeros((n + 31) // 32, np.uint32)
for name, fn in (("intersect (64 B in, 16 B out per ray)", lambda: e.Intersect(h)), ("occluded  (64 B in, 1 bit out)", lambda: e.IsOccluded(h, bits=bits))):
    fn()
    ts = []
    for _ in range(3):
        h["t"] = 1e30
        h["prim"] = 7
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
        if name.startswith("intersect"):
            assert (h["t"] == 5).all() and (h["prim"] == 0).all(), "host path returned wrong hits"
    t = min(ts)
    print(f"h2d {os.environ.get('TBVH_HOST_PATH', 'copy2d')} d2h_mode {os.environ.get('TBVH_D2H_MODE', '0')}: {name}: {t * 1e3:.1f} ms  {n / t / 1e6:.0f} Mrays/s  inbound {n * 64 / t / 1e9:.1f} GB/s")
if "--ref" not in sys.argv:
    sys.exit(0)
# reference points: plain 1D pinned copies
x = torch.empty(n * 64, dtype=torch.uint8).pin_memory()
d = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"1D pinned H2D {n * 64 / (t1 - t0) / 1e9:.1f} GB/s")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"1D pinned D2H {n * 64 / (t1 - t0) / 1e9:.1f} GB/s")
