"""ctypes binding of oracle/libtbvh_oracle.so (the plain-C restatement, oracle/tbvh_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference arms.  The product package (tinybvh_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libtbvh_oracle.so")
_lib = None

NODE32 = np.dtype([("aabbMin", "3f4"), ("leftFirst", "u4"), ("aabbMax", "3f4"), ("triCount", "u4")])
NODE64 = np.dtype([("lmin", "3f4"), ("left", "u4"), ("lmax", "3f4"), ("right", "u4"),
                   ("rmin", "3f4"), ("triCount", "u4"), ("rmax", "3f4"), ("firstTri", "u4")])


def build_lib(force: bool = False):
    srcs = [os.path.join(_HERE, f) for f in ("tbvh_oracle.c", "tbvh_oracle_hq.c", "tbvh_oracle_cwbvh.c", "tbvh_oracle.h")]
    if force or not os.path.isfile(PORT_SO) or any(os.path.getmtime(PORT_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(PORT_SO)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        L.orc_build.restype, L.orc_build.argtypes = u32, [vp, u32, vp, vp, f32, f32]
        L.orc_build_avx.restype, L.orc_build_avx.argtypes = u32, [vp, u32, vp, vp, f32, f32]
        L.orc_build_hq.restype, L.orc_build_hq.argtypes = u32, [vp, u32, vp, vp, vp, vp, f32, f32]
        L.orc_clip_frag.restype, L.orc_clip_frag.argtypes = i32, [vp, vp, vp, vp, vp, vp, u32]
        L.orc_split_frag.restype, L.orc_split_frag.argtypes = None, [vp, vp, vp, vp, vp, u32, f32, vp, vp]
        L.orc_intersect.restype, L.orc_intersect.argtypes = None, [vp, vp, vp, vp, u64, i32]
        L.orc_occluded.restype, L.orc_occluded.argtypes = None, [vp, vp, vp, vp, u64, vp, i32]
        L.orc_tri_test.restype, L.orc_tri_test.argtypes = i32, [vp] * 5 + [f32] + [vp] * 3
        L.orc_to_bvh_gpu.restype, L.orc_to_bvh_gpu.argtypes = u32, [vp, vp]
        L.orc_intersect_tlas.restype, L.orc_intersect_tlas.argtypes = None, [vp, vp, vp, vp, vp, u64]
        L.orc_occluded_tlas.restype, L.orc_occluded_tlas.argtypes = None, [vp, vp, vp, vp, vp, u64, vp]
        L.orc_instance_update.restype, L.orc_instance_update.argtypes = None, [vp, vp, vp]
        L.orc_intersect_tlas_cw.restype, L.orc_intersect_tlas_cw.argtypes = None, [vp, vp, vp, vp, vp, u64]
        L.orc_occluded_tlas_cw.restype, L.orc_occluded_tlas_cw.argtypes = None, [vp, vp, vp, vp, vp, u64, vp]
        L.orc_cwbvh_from_bvh.restype, L.orc_cwbvh_from_bvh.argtypes = u32, [vp, u32, vp, u32, vp, u32, vp, vp]
        L.orc_cwbvh_intersect.restype, L.orc_cwbvh_intersect.argtypes = None, [vp, vp, vp, u64]
        L.orc_refit.restype, L.orc_refit.argtypes = None, [vp, u32, vp, vp]
        L.orc_sah_cost.restype, L.orc_sah_cost.argtypes = f32, [vp, u32, f32, f32]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PortBVH:
    """orc_build + orc_intersect / orc_occluded: restatement of BVH::Build + Intersect / IsOccluded."""

    def __init__(self, verts=None, c_trav: float = 1.0, c_int: float = 1.0, nodes=None, prim_idx=None, avx: bool = False):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
        n = self.verts.shape[0] // 3
        if nodes is not None:
            self.nodes = np.ascontiguousarray(nodes).view(NODE32).reshape(-1)
            self.prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
            return
        nodes = np.zeros(max(2 * n, 2), NODE32)
        self.prim_idx = np.zeros(n, np.uint32)
        fn = lib().orc_build_avx if avx else lib().orc_build
        used = fn(_ptr(self.verts), n, _ptr(nodes), _ptr(self.prim_idx), c_trav, c_int)
        self.nodes = nodes[:used].copy()

    used_nodes = property(lambda s: s.nodes.shape[0])

    def intersect(self, rays, threads: int = 0):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        lib().orc_intersect(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.verts), _ptr(rays), rays.shape[0], threads)
        return rays

    def occluded(self, rays, threads: int = 0):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        bits = np.zeros((rays.shape[0] + 31) // 32, np.uint32)
        lib().orc_occluded(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.verts), _ptr(rays), rays.shape[0], _ptr(bits), threads)
        return bits

    def refit(self, new_verts):
        """orc_refit: BVH::Refit with the vertex array replaced by new_verts (same topology)."""
        self.verts = np.ascontiguousarray(new_verts, np.float32).reshape(-1, 4)
        self.nodes = self.nodes.copy()
        lib().orc_refit(_ptr(self.nodes), self.nodes.shape[0], _ptr(self.prim_idx), _ptr(self.verts))

    def to_bvh_gpu(self):
        out = np.zeros(self.nodes.shape[0], NODE64)
        used = lib().orc_to_bvh_gpu(_ptr(self.nodes), _ptr(out))
        return out[:used].copy()

    def sah_cost(self, c_trav=1.0, c_int=1.0):
        return float(lib().orc_sah_cost(_ptr(self.nodes), 0, c_trav, c_int))


class PortTLAS:
    """orc_intersect_tlas / orc_occluded_tlas: a TLAS (nodes + primIdx over instance boxes) walked with PortBVH BLASses.
    instances: the reference's 192-byte BLASInstance records, already updated (invTransform, box)."""

    def __init__(self, nodes, prim_idx, instances, blasses):
        self.nodes = np.ascontiguousarray(nodes).view(NODE32).reshape(-1)
        self.prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
        self.instances = np.ascontiguousarray(instances)
        assert self.instances.dtype.itemsize == 192
        self.blasses = list(blasses)
        class _B(C.Structure):
            _fields_ = [("nodes", C.c_void_p), ("primIdx", C.c_void_p), ("verts", C.c_void_p)]
        self._table = (_B * len(self.blasses))(*[_B(b.nodes.ctypes.data, b.prim_idx.ctypes.data, b.verts.ctypes.data) for b in self.blasses])

    def intersect(self, rays):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        lib().orc_intersect_tlas(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.instances), C.cast(self._table, C.c_void_p), _ptr(rays), rays.shape[0])
        return rays

    def occluded(self, rays):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        bits = np.zeros((rays.shape[0] + 31) // 32, np.uint32)
        lib().orc_occluded_tlas(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.instances), C.cast(self._table, C.c_void_p), _ptr(rays), rays.shape[0], _ptr(bits))
        return bits


class PortCWBVH:
    """orc_cwbvh_from_bvh + orc_cwbvh_intersect: the CWBVH conversion chain over a BVH2 (nodes, primIdx[, idxCount]) and the
    reference's CPU walk of the result."""

    def __init__(self, nodes, prim_idx, verts, idx_count=None):
        nodes = np.ascontiguousarray(nodes).view(NODE32).reshape(-1)
        prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
        n = self.verts.shape[0] // 3
        ic = int(idx_count) if idx_count is not None else prim_idx.shape[0]
        pidx = np.zeros(ic, np.uint32)
        pidx[: prim_idx.shape[0]] = prim_idx
        data = np.zeros((n * 5, 4), np.float32)
        self.tris = np.zeros((ic * 3, 4), np.float32)
        blocks = lib().orc_cwbvh_from_bvh(_ptr(nodes), nodes.shape[0], _ptr(pidx), ic, _ptr(self.verts), n, _ptr(data), _ptr(self.tris))
        self.nodes = data[:blocks].copy()

    def intersect(self, rays):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        lib().orc_cwbvh_intersect(_ptr(self.nodes), _ptr(self.tris), _ptr(rays), rays.shape[0])
        return rays


class PortTLASCW:
    """orc_intersect_tlas_cw / orc_occluded_tlas_cw: the TLAS walk of IntersectTLAS with BVH8_CWBVH::Intersect as the per-instance BLAS step
    (a composition of two pinned restatements - the reference's CPU TLAS does not take CWBVH BLASses; see tbvh_oracle.h).
    blasses: PortCWBVH objects (or anything with .nodes / .tris float4 arrays)."""

    def __init__(self, nodes, prim_idx, instances, blasses):
        self.nodes = np.ascontiguousarray(nodes).view(NODE32).reshape(-1)
        self.prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
        self.instances = np.ascontiguousarray(instances)
        assert self.instances.dtype.itemsize == 192
        self.blasses = [(np.ascontiguousarray(b.nodes, np.float32), np.ascontiguousarray(b.tris, np.float32)) for b in blasses]
        class _B(C.Structure):
            _fields_ = [("data", C.c_void_p), ("tris", C.c_void_p)]
        self._table = (_B * len(self.blasses))(*[_B(d.ctypes.data, t.ctypes.data) for d, t in self.blasses])

    def intersect(self, rays):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        lib().orc_intersect_tlas_cw(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.instances), C.cast(self._table, C.c_void_p), _ptr(rays), rays.shape[0])
        return rays

    def occluded(self, rays):
        assert rays.dtype.itemsize == 128 and rays.flags.c_contiguous
        bits = np.zeros((rays.shape[0] + 31) // 32, np.uint32)
        lib().orc_occluded_tlas_cw(_ptr(self.nodes), _ptr(self.prim_idx), _ptr(self.instances), C.cast(self._table, C.c_void_p), _ptr(rays), rays.shape[0], _ptr(bits))
        return bits


def instance_update(instances, bmin, bmax):
    """orc_instance_update on every record of a BLASInstance array (in place), for BLAS root boxes bmin[i] / bmax[i] (or one box)."""
    lo, hi = np.ascontiguousarray(bmin, np.float32).reshape(-1, 3), np.ascontiguousarray(bmax, np.float32).reshape(-1, 3)
    for i in range(instances.shape[0]):
        a, b = lo[i % lo.shape[0]].copy(), hi[i % hi.shape[0]].copy()
        lib().orc_instance_update(instances[i:i + 1].ctypes.data, _ptr(a), _ptr(b))
    return instances


def build_hq(verts, c_trav: float = 1.0, c_int: float = 1.0):
    """orc_build_hq: restatement of BVH::BuildHQ + Compact -> (nodes, primIdx[:usedIdx], idxCount)."""
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 4)
    n = v.shape[0] // 3
    nodes = np.zeros(3 * n + 2, NODE32)
    idx = np.zeros(n + n // 2 + 1, np.uint32)
    ic, ui = C.c_uint32(), C.c_uint32()
    used = lib().orc_build_hq(_ptr(v), n, _ptr(nodes), _ptr(idx), C.byref(ic), C.byref(ui), c_trav, c_int)
    return nodes[:used].copy(), idx[:ui.value].copy(), ic.value


FRAGMENT = np.dtype([("bmin", "3f4"), ("primIdx", "u4"), ("bmax", "3f4"), ("clipped", "u4")])


def clip_frag(verts, frag, bmin, bmax, min_dim, axis):
    out = np.zeros(1, FRAGMENT)
    a = [np.ascontiguousarray(x, np.float32) for x in (bmin, bmax, min_dim)]
    ok = lib().orc_clip_frag(_ptr(verts), _ptr(frag), _ptr(out), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), int(axis))
    return bool(ok), out


def split_frag(verts, frag, min_dim, axis, pos):
    l, r = np.zeros(1, FRAGMENT), np.zeros(1, FRAGMENT)
    lo, ro = C.c_int(), C.c_int()
    md = np.ascontiguousarray(min_dim, np.float32)
    lib().orc_split_frag(_ptr(verts), _ptr(frag), _ptr(l), _ptr(r), _ptr(md), int(axis), float(pos), C.byref(lo), C.byref(ro))
    return bool(lo.value), bool(ro.value), l, r


def tri_test(O, D, v0, v1, v2, tmax):
    a = [np.ascontiguousarray(x, np.float32) for x in (O, D, v0, v1, v2)]
    t, u, v = (np.zeros(1, np.float32) for _ in range(3))
    ok = lib().orc_tri_test(*[_ptr(x) for x in a], float(tmax), _ptr(t), _ptr(u), _ptr(v))
    return bool(ok), float(t[0]), float(u[0]), float(v[0])
