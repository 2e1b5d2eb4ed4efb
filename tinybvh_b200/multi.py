"""Multi-GPU plumbing for the hot path: rays shard by index, the BVH is built once and broadcast.

One process per GPU (torch.distributed).  Traversal needs no data-path collective: every rank holds a full replica of
the BVH and traces a contiguous slice of the ray batch; the only exchange is ONE broadcast of the node / index / vertex
arrays from the building rank over NVLink (NCCL; gloo on CPU for the host-logic tests).  Results are disjoint, so there
is no reduction; occlusion bit masks are sliced on 32-ray boundaries so no word is shared between ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, align: int = 32):
    """Contiguous slice [start, start+count) of n rays for `rank`; boundaries are multiples of `align` so occlusion
    words never straddle ranks.  Slices cover [0,n) exactly once."""
    units = (n + align - 1) // align
    lo = (units * rank) // world
    hi = (units * (rank + 1)) // world
    start, end = min(lo * align, n), min(hi * align, n)
    return start, end - start


def block_cyclic(n: int, rank: int, world: int, block: int = 1 << 20):
    """The rays of `rank` when a set of n rays is dealt out in blocks of `block` consecutive rays, block b to rank b % world: a list of
    (first, count).  Blocks keep 4x4-pixel tiles and warps coherent, dealing them round-robin keeps every rank's share of the image
    representative (contiguous eighths of a view differ by tens of per cent in traversal work).  `block` must be a multiple of 32 so
    occlusion words never straddle two ranks.  world == 1 gives the whole set as one block."""
    assert block % 32 == 0 and 0 <= rank < world
    if world == 1:
        return [(0, n)] if n else []
    return [(b * block, min(block, n - b * block)) for b in range((n + block - 1) // block) if b % world == rank]


def broadcast_arrays(arrays, src: int = 0, device=None):
    """Broadcast a dict of tensors from `src` to every rank (one dist.broadcast per tensor after a metadata round).
    On non-src ranks `arrays` may be None.  dtypes are restricted to int32 / float32 / uint8 for portability."""
    rank = dist.get_rank()
    dev = device if device is not None else (next(iter(arrays.values())).device if arrays else torch.device("cpu"))
    names = ["nodes", "prim_idx", "verts"]
    codes = {torch.int32: 0, torch.float32: 1, torch.uint8: 2}
    inv = {v: k for k, v in codes.items()}
    meta = torch.zeros(len(names) * 2, dtype=torch.int64, device=dev)
    if rank == src:
        for i, k in enumerate(names):
            meta[2 * i], meta[2 * i + 1] = arrays[k].numel(), codes[arrays[k].dtype]
    dist.broadcast(meta, src)
    out = {}
    m = meta.cpu().tolist()
    for i, k in enumerate(names):
        t = arrays[k].contiguous() if rank == src else torch.empty(int(m[2 * i]), dtype=inv[int(m[2 * i + 1])], device=dev)
        dist.broadcast(t, src)
        out[k] = t
    return out
