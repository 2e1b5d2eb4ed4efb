"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol the header declares;
a data call without a CUDA device fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tinybvh_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    so = build.build()
    assert os.path.isfile(so)
    L = ctypes.CDLL(so)
    assert L is not None


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(REPO, "include", "tinybvh_b200.h")).read()
    declared = set(re.findall(r"\b(tbvh_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tbvh_ctx_t", "tbvh_bvh_t"}
    assert len(declared) >= 20
    L = ctypes.CDLL(build.build())
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"header declares symbols the library does not export: {missing}"
    # and the Python binding table covers exactly the header
    assert set(_lib.SYMBOLS) == declared


def test_header_cites_reference_lines():
    hdr = open(os.path.join(REPO, "include", "tinybvh_b200.h")).read()
    assert hdr.count("tiny_bvh.h:") >= 10


def test_no_cpu_fallback_without_device():
    L = _lib.lib()
    if L.tbvh_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = L.tbvh_ctx_create(0, ctypes.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in L.tbvh_last_error() or b"cuda" in L.tbvh_last_error().lower()


def test_instance_update_matches_the_oracle():
    """tbvh_instance_update_box is host arithmetic (no device needed): BLASInstance::Update (tiny_bvh.h:8386), bit for bit."""
    import ctypes as C
    import numpy as np
    from oracle import portpy
    from tinybvh_b200 import api
    rng = np.random.default_rng(3)
    T = (rng.random((500, 16), np.float32) - 0.5) * 4
    T[::2, 12:15], T[::2, 15] = 0, 1
    a = np.zeros(500, api.BLAS_INSTANCE)
    a["transform"] = T
    a["mask"] = 0xFFFF
    b = a.copy()
    lo, hi = np.array([-1.5, -0.7, -2.2], np.float32), np.array([1.1, 2.3, 0.9], np.float32)
    portpy.instance_update(a, lo, hi)
    L = _lib.lib()
    for i in range(500):
        assert L.tbvh_instance_update_box(C.c_void_p(b[i:i + 1].ctypes.data), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p)) == 0
    assert a.tobytes() == b.tobytes()


def test_sah_cost_of_a_host_node_array_matches_the_oracle():
    """tbvh_sah_cost_nodes is host arithmetic: BVH::SAHCost (tiny_bvh.h:1889) bit for bit, on golden trees incl. the SBVH ones."""
    import ctypes as C
    import numpy as np
    from oracle import portpy
    from tests import golden_util as G
    L = _lib.lib()
    for path in G.golden_files():
        g = G.load(path)
        for key in ("nodes", "hq_nodes"):
            nodes = np.ascontiguousarray(g[key])
            want = np.float32(portpy.lib().orc_sah_cost(nodes.ctypes.data_as(C.c_void_p), 0, 1.0, 1.0))
            out = C.c_float()
            assert L.tbvh_sah_cost_nodes(nodes.ctypes.data_as(C.c_void_p), nodes.shape[0], 1.0, 1.0, C.byref(out)) == 0
            assert np.float32(out.value).view(np.uint32) == want.view(np.uint32), (path, key)
