"""Triangle-soup scenes for the tinybvh hot path.

File format (reference loader: tiny_bvh_speedtest.cpp:486-495): uint32 triCount, then triCount*3 float4
vertices (16-byte stride, w = packed colour / unused).  The reference's fixtures live in
/root/reference/testdata; that path does not exist on the GPU box, so `tools/fetch_scenes.py` copies the
ones the configs name into data/scenes/ (git-ignored, travels with the gpurun snapshot).  When a fixture is
missing, `load_scene` falls back to a deterministic procedural scene with the same triangle count and says
so in the returned name ("synthetic:...") - bench.py reports that in its `data` field.
"""
from __future__ import annotations

import os
import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIRS = [os.path.join(_REPO, "data", "scenes"), "/root/reference/testdata"]

# name -> (files concatenated in order, triangle count) ; counts verified in SURVEY.md section 4
SCENES = {
    "bunny": (["bunny.bin"], 69630),
    "sponza": (["cryteksponza.bin"], 262267),
    "bistro": (["bistro_ext_part1.bin", "bistro_ext_part2.bin"], 2837209),
    "lucy_dragon": (["lucy.bin", "xyzrgb_dragon.bin"], 349852),
    "legocar": (["legocar.bin"], 10992),
    "suzanne": (["suzanne.bin"], 15488),
    "head": (["head.bin"], 17684),
}


def read_bin(path: str) -> np.ndarray:
    """-> float32 [ntris*3, 4] vertex array."""
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
        v = np.frombuffer(f.read(n * 48), dtype=np.float32)
    if v.size != n * 12:
        raise IOError(f"{path}: truncated ({v.size} floats for {n} tris)")
    return v.reshape(n * 3, 4).copy()


def write_bin(path: str, verts: np.ndarray) -> None:
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 4)
    with open(path, "wb") as f:
        f.write(np.uint32(verts.shape[0] // 3).tobytes())
        f.write(verts.tobytes())


def _find(fname: str):
    for d in SCENE_DIRS:
        p = os.path.join(d, fname)
        if os.path.isfile(p):
            return p
    return None


def procedural_scene(ntris: int, seed: int = 1234) -> np.ndarray:
    """Deterministic 'atrium' stand-in: a box room with tessellated floor / walls, rows of faceted columns
    and a cloud of small displaced blobs, roughly Sponza-like in extent (~[-40,40] x [0,30] x [-20,20]) and in
    its mix of large and tiny triangles.  Exactly `ntris` triangles."""
    rng = np.random.default_rng(seed)
    tris = []

    def quad_grid(origin, du, dv, nu, nv, jitter=0.0):
        o = np.asarray(origin, np.float64)
        du = np.asarray(du, np.float64)
        dv = np.asarray(dv, np.float64)
        i, j = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="ij")
        p = o + i[..., None] * du / nu + j[..., None] * dv / nv
        if jitter:
            n = np.cross(du, dv)
            n /= np.linalg.norm(n)
            p = p + n * (rng.random(p.shape[:2])[..., None] - 0.5) * jitter
        a, b, c, d = p[:-1, :-1], p[1:, :-1], p[1:, 1:], p[:-1, 1:]
        t = np.concatenate([np.stack([a, b, c], 2).reshape(-1, 3, 3), np.stack([a, c, d], 2).reshape(-1, 3, 3)])
        tris.append(t)

    def column(cx, cz, r, h, seg, rings):
        ang = np.linspace(0, 2 * np.pi, seg + 1)
        ys = np.linspace(0, h, rings + 1)
        rr = r * (1 + 0.08 * np.sin(ys * 3.0))
        p = np.stack([cx + np.outer(rr, np.cos(ang)), np.repeat(ys[:, None], seg + 1, 1), cz + np.outer(rr, np.sin(ang))], -1)
        a, b, c, d = p[:-1, :-1], p[1:, :-1], p[1:, 1:], p[:-1, 1:]
        tris.append(np.concatenate([np.stack([a, b, c], 2).reshape(-1, 3, 3), np.stack([a, c, d], 2).reshape(-1, 3, 3)]))

    budget = ntris
    g = max(4, int(np.sqrt(budget * 0.10 / 2 / 5)))
    quad_grid((-40, 0, -20), (80, 0, 0), (0, 0, 40), 2 * g, g, 0.05)     # floor
    quad_grid((-40, 30, -20), (80, 0, 0), (0, 0, 40), g, g // 2 + 1)     # ceiling
    quad_grid((-40, 0, -20), (80, 0, 0), (0, 30, 0), g, g // 2 + 1, 0.1)  # walls
    quad_grid((-40, 0, 20), (80, 0, 0), (0, 30, 0), g, g // 2 + 1, 0.1)
    quad_grid((-40, 0, -20), (0, 0, 40), (0, 30, 0), g // 2 + 1, g // 2 + 1)
    quad_grid((40, 0, -20), (0, 0, 40), (0, 30, 0), g // 2 + 1, g // 2 + 1)
    ncol = 24
    seg = max(6, int(np.sqrt(budget * 0.25 / ncol / 2)))
    for k in range(ncol):
        column(-33 + 6 * (k % 12), -9 if k < 12 else 9, 1.2, 18, seg, seg)
    have = sum(t.shape[0] for t in tris)
    # blobs: small icosphere-ish shells made of random thin triangles around random centres
    rest = max(0, budget - have)
    if rest:
        nb = max(1, rest // 400)
        cen = np.stack([rng.uniform(-36, 36, nb), rng.uniform(0.5, 25, nb), rng.uniform(-17, 17, nb)], -1)
        rad = rng.uniform(0.2, 1.5, nb)
        owner = rng.integers(0, nb, rest)
        d0 = rng.normal(size=(rest, 3))
        d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
        t1 = np.cross(d0, rng.normal(size=(rest, 3)))
        t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
        t2 = np.cross(d0, t1)
        base = cen[owner] + d0 * rad[owner, None]
        s = (rad[owner] * rng.uniform(0.05, 0.25, rest))[:, None]
        tris.append(np.stack([base, base + t1 * s, base + t2 * s], 1))
    t = np.concatenate(tris)[:ntris]
    if t.shape[0] < ntris:  # top up by repeating shifted blobs
        extra = t[rng.integers(0, t.shape[0], ntris - t.shape[0])] + rng.normal(scale=0.01, size=(ntris - t.shape[0], 1, 3))
        t = np.concatenate([t, extra])
    out = np.zeros((ntris * 3, 4), np.float32)
    out[:, :3] = t.reshape(-1, 3).astype(np.float32)
    return out


def replicate_grid(verts: np.ndarray, copies: int, grid=(4, 4, 2), pitch: float = 1.1) -> np.ndarray:
    """replicate a mesh `copies` times on a grid with cell pitch = pitch * bbox extent."""
    v = verts.reshape(-1, 4)
    lo, hi = v[:, :3].min(0), v[:, :3].max(0)
    ext = (hi - lo) * pitch
    out = []
    k = 0
    for z in range(grid[2]):
        for y in range(grid[1]):
            for x in range(grid[0]):
                if k >= copies:
                    break
                w = v.copy()
                w[:, :3] += (np.array([x, y, z], np.float32) * ext).astype(np.float32)
                out.append(w)
                k += 1
    return np.concatenate(out)


def load_scene(name: str, allow_synthetic: bool = True):
    """-> (verts float32 [ntris*3,4], label).  label is the scene name, or 'synthetic:<name>' for the fallback."""
    if name.startswith("synthetic:"):
        n = int(name.split(":")[1])
        return procedural_scene(n), name
    if name == "lucy_dragon_x29":
        # BASELINE.json configs[4] says "~10M tris combined"; the shipped fixtures total 349,852, so the mesh is replicated
        # 29x on a 4x4x2 grid with pitch 1.1 x bbox extent -> 10,145,708 triangles
        base, label = load_scene("lucy_dragon", allow_synthetic)
        return replicate_grid(base, 29), ("lucy_dragon_x29" if label == "lucy_dragon" else "synthetic:lucy_dragon_x29")
    files, ntris = SCENES[name]
    paths = [_find(f) for f in files]
    if all(paths):
        v = np.concatenate([read_bin(p) for p in paths])
        if v.shape[0] != ntris * 3:
            raise IOError(f"scene {name}: expected {ntris} tris, files hold {v.shape[0] // 3}")
        return v, name
    if not allow_synthetic:
        raise FileNotFoundError(f"scene {name}: {files} not found in {SCENE_DIRS}")
    return procedural_scene(ntris), f"synthetic:{name}"


def scene_bounds(verts: np.ndarray):
    v = verts.reshape(-1, 4)[:, :3]
    return v.min(0), v.max(0)
