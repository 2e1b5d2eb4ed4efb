#!/usr/bin/env python
"""BuildHQ timing on the GPU box: scenes x repeats, device milliseconds (info.build_ms) and tree statistics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinybvh_b200 import api, scenes

for sc in sys.argv[1:] or ["bunny", "sponza", "synthetic:1000000"]:
    try:
        v, label = scenes.load_scene(sc)
    except Exception as ex:
        print(sc, "unavailable:", ex); continue
    ms = []
    for k in range(4):
        t = time.time(); e = api.BVH().BuildHQ(v); wall = time.time() - t
        ms.append(e.info().build_ms)
    i = e.info()
    print(f"{label}: tris {v.shape[0] // 3} nodes {i.used_nodes} idxCount {i.idx_count} depth {i.max_depth} build_ms {['%.2f' % m for m in ms]} wall {wall * 1e3:.1f} ms", flush=True)
