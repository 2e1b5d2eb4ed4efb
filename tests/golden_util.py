import glob
import os

import numpy as np

from tinybvh_b200 import rays as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load(path):
    return dict(np.load(path))


def rays_of(g, kind):
    """Rebuild the 128-byte host records from the stored O, D, rD, tmax (stored rD is canonical: Appendix A.1)."""
    n = g[kind + "_O"].shape[0]
    r = np.zeros(n, R.RAY_DTYPE)
    r["O"], r["D"], r["rD"], r["t"] = g[kind + "_O"], g[kind + "_D"], g[kind + "_rD"], g[kind + "_tmax"]
    r["mask"] = 0xFFFF
    return r


def hits_as_u32(r):
    return np.stack([r["t"].view(np.uint32), r["u"].view(np.uint32), r["v"].view(np.uint32), r["prim"]], 1)
