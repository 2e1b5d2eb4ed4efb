"""Host ray records and the deterministic ray-set generators of SURVEY.md 8(d).

The host `Ray` is the reference's 128-byte, 64-aligned record (tiny_bvh.h:688-709); with either value of
INST_IDX_BITS the fields the GPU consumes sit at O=0, mask=12, D=16, instIdx=28, rD=32, t=48, u=52, v=56,
prim=60.  The GPU record is the first 64 bytes (traverse.cl:11-17, tiny_bvh_speedtest.cpp:1110-1115).
All arithmetic is float32 in the operation order of `Ray::Ray` (:695-701), `tinybvh_normalize` (:508-512)
and `tinybvh_safercp` (:442).
"""
from __future__ import annotations

import numpy as np

BVH_FAR = np.float32(1e30)

RAY_DTYPE = np.dtype({
    "names": ["O", "mask", "D", "instIdx", "rD", "pad", "t", "u", "v", "prim", "aux"],
    "formats": ["3f4", "u4", "3f4", "u4", "3f4", "u4", "f4", "f4", "f4", "u4", ("u1", 64)],
    "offsets": [0, 12, 16, 28, 32, 44, 48, 52, 56, 60, 64],
    "itemsize": 128,
})
# the 64-byte device record, and the 16-byte hit record (t, u, v, prim)
GPU_RAY_DTYPE = np.dtype({
    "names": ["O", "mask", "D", "instIdx", "rD", "pad", "t", "u", "v", "prim"],
    "formats": ["3f4", "u4", "3f4", "u4", "3f4", "u4", "f4", "f4", "f4", "u4"],
    "offsets": [0, 12, 16, 28, 32, 44, 48, 52, 56, 60],
    "itemsize": 64,
})
HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("prim", "u4")])

f32 = np.float32


def normalize(v: np.ndarray) -> np.ndarray:
    """tinybvh_normalize: l = sqrtf(x*x+y*y+z*z); rl = l==0 ? 0 : 1/l; v*rl   (float32 throughout)."""
    v = np.asarray(v, f32)
    l = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2], dtype=f32)
    with np.errstate(divide="ignore"):
        rl = np.where(l == 0, f32(0), f32(1) / l).astype(f32)
    return (v * rl[..., None]).astype(f32)


def safercp(d: np.ndarray) -> np.ndarray:
    """tinybvh_safercp (:442): 1/x when |x| > 1e-12, else +-1e30 by the sign test `x >= 0`."""
    d = np.asarray(d, f32)
    big = (d > f32(1e-12)) | (d < f32(-1e-12))
    with np.errstate(divide="ignore"):
        r = f32(1) / np.where(big, d, f32(1))
    return np.where(big, r, np.where(d >= 0, BVH_FAR, -BVH_FAR)).astype(f32)


def make_rays(O: np.ndarray, D: np.ndarray, tmax=BVH_FAR, normalized: bool = False) -> np.ndarray:
    """Ray::Ray for a batch: zeroed 128-byte records, D normalised, rD = safercp(D), hit.t = tmax, mask 0xFFFF."""
    O = np.asarray(O, f32).reshape(-1, 3)
    D = np.asarray(D, f32).reshape(-1, 3)
    r = np.zeros(O.shape[0], RAY_DTYPE)
    r["O"] = O
    r["D"] = D if normalized else normalize(D)
    r["rD"] = safercp(r["D"])
    r["t"] = tmax
    r["mask"] = 0xFFFF
    return r


def reset_hits(rays: np.ndarray, tmax=BVH_FAR) -> None:
    rays["t"] = tmax
    rays["u"] = 0
    rays["v"] = 0
    rays["prim"] = 0


# ---------------------------------------------------------------------------------------------- cameras
SPONZA_EYES = np.array([[-15.24, 21.5, 2.54], [-34, 5, 11.26], [-1.3, 4.96, 12.28]], f32)
SPONZA_VIEWS = np.array([[0.826, -0.438, -0.356], [0.9427, 0.0292, -0.3324], [-0.9886, 0.0507, -0.1419]], f32)


def bounds_camera(lo, hi, kind: str = "outside"):
    """Cameras for scenes the reference ships no camera for.  'outside' = SURVEY config 1 (eye = bbox centre +
    (0,0,2*maxExtent) looking at the centre); 'inside' = the survey's Bistro probe camera."""
    lo, hi = np.asarray(lo, f32), np.asarray(hi, f32)
    c, e = (lo + hi) * f32(0.5), hi - lo
    if kind == "outside":
        eye = c + np.array([0, 0, 2 * e.max()], f32)
        view = c - eye
    else:
        eye = c + np.array([-0.2 * e[0], 0.1 * e[1], 0.05 * e[2]], f32)
        view = np.array([1, -0.05, 0.2], f32)
    return eye.astype(f32), normalize(view)


def primary_rays(eye, view, width: int, height: int, spp: int = 16) -> np.ndarray:
    """Pinhole pattern of tiny_bvh_speedtest.cpp:497-551: 4x4-pixel tiles, 16 samples per pixel on a 4x4
    sub-grid (u=(4px+(s&3))/(4W), v=(4py+(s>>2))/(4H)); spp<16 keeps samples with (s % (16/spp)) == 0
    (spp=2 is the speedtest's 'smallBatch')."""
    assert width % 4 == 0 and height % 4 == 0 and 16 % spp == 0
    eye = np.asarray(eye, f32)
    view = normalize(np.asarray(view, f32))
    right = normalize(np.cross(np.array([0, 1, 0], f32), view).astype(f32))
    up = (f32(0.8) * np.cross(view, right)).astype(f32)
    C = (eye + f32(2) * view).astype(f32)
    p1, p2, p3 = C - right + up, C + right + up, C - right - up
    ty, tx, y, x, s = np.meshgrid(np.arange(height // 4), np.arange(width // 4), np.arange(4), np.arange(4),
                                  np.arange(0, 16, 16 // spp), indexing="ij")
    px, py = (tx * 4 + x).ravel(), (ty * 4 + y).ravel()
    s = s.ravel()
    u = ((px * 4 + (s & 3)).astype(f32) / f32(width * 4)).astype(f32)
    v = ((py * 4 + (s >> 2)).astype(f32) / f32(height * 4)).astype(f32)
    P = (p1 + u[:, None] * (p2 - p1) + v[:, None] * (p3 - p1)).astype(f32)
    D = normalize(P - eye)
    return make_rays(np.broadcast_to(eye, D.shape), D)


def shadow_rays(primary: np.ndarray, light, eps: float) -> np.ndarray:
    """tiny_bvh_speedtest.cpp:853-865: from each primary hit point (t clamped to 1000) towards `light`,
    origin offset eps along D, tmax = |L-I| - eps.  `primary` must hold traced hits."""
    t = np.minimum(f32(1000), primary["t"]).astype(f32)
    I = (primary["O"] + t[:, None] * primary["D"]).astype(f32)
    L = np.asarray(light, f32)
    toL = (L - I).astype(f32)
    D = normalize(toL)
    dist = np.sqrt(toL[:, 0] * toL[:, 0] + toL[:, 1] * toL[:, 1] + toL[:, 2] * toL[:, 2], dtype=f32)
    return make_rays(I + D * f32(eps), D, (dist - f32(eps)).astype(f32), normalized=True)


def _xorshift(s):
    s ^= (s << np.uint32(13)) & np.uint32(0xFFFFFFFF)
    s ^= s >> np.uint32(17)
    s ^= (s << np.uint32(5)) & np.uint32(0xFFFFFFFF)
    return s


def _wang(seed):
    s = (seed ^ np.uint32(61)) ^ (seed >> np.uint32(16))
    s = (s * np.uint32(9)) & np.uint32(0xFFFFFFFF)
    s ^= s >> np.uint32(4)
    s = (s * np.uint32(0x27D4EB2D)) & np.uint32(0xFFFFFFFF)
    s ^= s >> np.uint32(15)
    return s


def diffuse_rays(primary: np.ndarray, verts: np.ndarray, seed: int = 0x123456) -> np.ndarray:
    """Incoherent bounce rays as tiny_bvh_speedtest.cpp:564-587, with `rand()` replaced by xorshift32
    (tiny_bvh.h:1549) seeded per ray with a Wang hash of the ray index: R = random unit
    vector flipped into the hemisphere of the geometric normal facing the ray, origin I + 0.001 R."""
    n = primary.shape[0]
    with np.errstate(over="ignore"):
        s = _wang((np.arange(n, dtype=np.uint32) + np.uint32(seed)) | np.uint32(1))
        R = np.empty((n, 3), f32)
        for k in range(3):
            s = _xorshift(s)
            R[:, k] = (s.astype(np.float64) * 2.3283064365387e-10).astype(f32) - f32(0.5)
    R = normalize(R)
    O, D, t = primary["O"], primary["D"], primary["t"]
    hit = t < f32(100)
    I = np.where(hit[:, None], O + t[:, None] * D, O + f32(20) * D).astype(f32)
    v = verts.reshape(-1, 3, 4)[:, :, :3]
    p = np.where(hit, primary["prim"], 0).astype(np.int64)
    N = normalize(np.cross(v[p, 1] - v[p, 0], v[p, 2] - v[p, 0]).astype(f32))
    N = np.where((np.einsum("ij,ij->i", N, D) > 0)[:, None], -N, N)
    flip = hit & (np.einsum("ij,ij->i", N, R) < 0)
    R = np.where(flip[:, None], -R, R).astype(f32)
    return make_rays(I + f32(0.001) * R, R, normalized=True)


def gpu_records(rays: np.ndarray) -> np.ndarray:
    """First 64 bytes of every 128-byte host record, contiguous (what tiny_bvh_speedtest.cpp:1110-1115 uploads)."""
    return np.ascontiguousarray(rays.view(np.uint8).reshape(-1, 128)[:, :64]).view(GPU_RAY_DTYPE).reshape(-1)


# ---------------------------------------------------------------------------------------------- fast generators (hostgen/raygen.c)
_gen = None


def _genlib():
    """libtbvh_raygen.so: the same generators in C + OpenMP (bit-identical records, tests/test_raygen.py); None when not built."""
    global _gen
    if _gen is None:
        import ctypes as C
        import os
        so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtbvh_raygen.so")
        if not os.path.isfile(so):
            _gen = False
            return None
        L = C.CDLL(so)
        vp, u64, u32, f = C.c_void_p, C.c_uint64, C.c_uint32, C.c_float
        L.tbvh_gen_primary.argtypes = [vp, u64, u64, vp, vp, vp, vp, u32, u32, u32]
        L.tbvh_gen_shadow.argtypes = [vp, vp, vp, u64, vp, f]
        L.tbvh_gen_diffuse.argtypes = [vp, vp, vp, u64, u64, vp, u32]
        L.tbvh_gen_reset_hits.argtypes = [vp, u64, f]
        L.tbvh_gen_store_hits.argtypes = [vp, vp, u64]
        L.tbvh_gen_set_threads.argtypes = [C.c_int]
        for fn in (L.tbvh_gen_primary, L.tbvh_gen_shadow, L.tbvh_gen_diffuse, L.tbvh_gen_reset_hits, L.tbvh_gen_store_hits, L.tbvh_gen_set_threads):
            fn.restype = None
        _gen = L
    return _gen or None


def set_generator_threads(n: int) -> None:
    """Threads the C generators may use (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    L = _genlib()
    if L is not None:
        L.tbvh_gen_set_threads(int(n))


def _p(a):
    return a.ctypes.data


def camera_plane(eye, view):
    """eye, p1, p2-p1, p3-p1 of the speedtest's view plane (tiny_bvh_speedtest.cpp:499-520), float32."""
    eye = np.asarray(eye, f32)
    view = normalize(np.asarray(view, f32))
    right = normalize(np.cross(np.array([0, 1, 0], f32), view).astype(f32))
    up = (f32(0.8) * np.cross(view, right)).astype(f32)
    C = (eye + f32(2) * view).astype(f32)
    p1, p2, p3 = C - right + up, C + right + up, C - right - up
    return eye, p1.astype(f32), (p2 - p1).astype(f32), (p3 - p1).astype(f32)


def primary_rays_into(out: np.ndarray, eye, view, width: int, height: int, spp: int = 16, first: int = 0) -> np.ndarray:
    """Fill `out` (RAY_DTYPE, e.g. page-locked) with rays [first, first+len(out)) of the camera set `primary_rays` defines."""
    assert out.dtype.itemsize == 128 and out.flags.c_contiguous and width % 4 == 0 and height % 4 == 0 and 16 % spp == 0
    n = out.shape[0]
    assert first + n <= width * height * spp
    L = _genlib()
    if L is None:
        out[:] = primary_rays(eye, view, width, height, spp)[first:first + n]
        return out
    e, p1, d21, d31 = [np.ascontiguousarray(x, f32) for x in camera_plane(eye, view)]
    L.tbvh_gen_primary(_p(out), first, n, _p(e), _p(p1), _p(d21), _p(d31), width, height, spp)
    return out


def shadow_rays_into(out: np.ndarray, primary: np.ndarray, light, eps: float, hits: np.ndarray = None) -> np.ndarray:
    """`shadow_rays` into `out`; `hits` (packed t,u,v,prim per ray) replaces the records' own hit.t when given."""
    assert out.dtype.itemsize == 128 and primary.dtype.itemsize == 128 and out.shape[0] == primary.shape[0]
    L = _genlib()
    if L is None:
        src = primary
        if hits is not None:
            src = primary.copy()
            src["t"] = np.asarray(hits).reshape(-1, 4)[:, 0]
        out[:] = shadow_rays(src, light, eps)
        return out
    light = np.ascontiguousarray(light, f32)
    hp = None
    if hits is not None:
        hits = np.ascontiguousarray(hits).view(f32).reshape(-1, 4)
        assert hits.shape[0] == primary.shape[0]
        hp = _p(hits)
    L.tbvh_gen_shadow(_p(out), _p(primary), hp, out.shape[0], _p(light), np.float32(eps))
    return out


def diffuse_rays_into(out: np.ndarray, primary: np.ndarray, verts: np.ndarray, hits: np.ndarray = None, seed: int = 0x123456, first: int = 0) -> np.ndarray:
    """`diffuse_rays` into `out` for rays [first, first+len(out)) of the whole set (the per-ray seed is the global ray index)."""
    assert out.dtype.itemsize == 128 and primary.dtype.itemsize == 128 and out.shape[0] == primary.shape[0]
    L = _genlib()
    if L is None:
        assert first == 0
        src = primary
        if hits is not None:
            src = primary.copy()
            h = np.asarray(hits).view(f32).reshape(-1, 4)
            src["t"], src["prim"] = h[:, 0], h[:, 3].view(np.uint32)
        out[:] = diffuse_rays(src, verts, seed)
        return out
    verts = np.ascontiguousarray(verts, f32)
    hp = None
    if hits is not None:
        hits = np.ascontiguousarray(hits).view(f32).reshape(-1, 4)
        hp = _p(hits)
    L.tbvh_gen_diffuse(_p(out), _p(primary), hp, first, out.shape[0], _p(verts), seed)
    return out


def reset_hits_fast(rays: np.ndarray, tmax=BVH_FAR) -> None:
    L = _genlib()
    if L is None or rays.dtype.itemsize != 128:
        return reset_hits(rays, tmax)
    L.tbvh_gen_reset_hits(_p(rays), rays.shape[0], np.float32(tmax))
