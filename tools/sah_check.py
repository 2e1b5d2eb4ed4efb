#!/usr/bin/env python
"""GPU-box check: BVH.SAHCost() of GPU-built trees (Build, BuildAVX, BuildHQ) against the reference's SAHCost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinybvh_b200 import api, scenes
from oracle import refpy
ok = True
for sc in ("bunny", "sponza"):
    v, label = scenes.load_scene(sc)
    for mode, name in ((0, "Build"), (1, "BuildAVX"), (2, "BuildHQ")):
        want = np.float32(refpy.RefBVH(v, mode=mode, threaded=False).sah_cost())
        got = np.float32(getattr(api.BVH(), name)(v).SAHCost())
        same = want.view(np.uint32) == got.view(np.uint32)
        ok &= bool(same)
        print(f"{label} {name}: SAH {got} (reference {want}) {'identical' if same else 'DIFFERS'}")
sys.exit(0 if ok else 1)
