"""N>1 host logic on CPU: world_size-2 gloo run of the BVH broadcast + ray sharding plumbing (tinybvh_b200/multi.py).
Each rank receives rank 0's tree, traces ITS shard with the oracle port standing in for the GPU engine (test only), and
rank 0 checks that the concatenated shards equal a single-process trace of the whole batch."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from tinybvh_b200 import multi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_exactly_once_and_align():
    for n in (0, 1, 31, 32, 33, 1000, 16777216, 12345677):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                s, c = multi.shard_range(n, r, world)
                assert s == pos and (s % 32 == 0 or s == n)
                pos += c
            assert pos == n


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.environ["TBVH_REPO"])
    import numpy as np, torch, torch.distributed as dist
    from tinybvh_b200 import multi, scenes
    from oracle import portpy
    from tests import util
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    v = scenes.procedural_scene(3000, 41)
    sets, _ = util.ray_sets(v, res=48)
    rays = sets["primary"]
    arrays = None
    if rank == 0:
        p = portpy.PortBVH(v)
        arrays = {"nodes": torch.from_numpy(p.nodes.view(np.int32).copy()), "prim_idx": torch.from_numpy(p.prim_idx.view(np.int32).copy()),
                  "verts": torch.from_numpy(v.reshape(-1).copy())}
    got = multi.broadcast_arrays(arrays, 0, torch.device("cpu"))
    nodes = got["nodes"].numpy().view(portpy.NODE32).reshape(-1)
    local = portpy.PortBVH(got["verts"].numpy().reshape(-1, 4), nodes=nodes, prim_idx=got["prim_idx"].numpy().view(np.uint32))
    s, c = multi.shard_range(rays.shape[0], rank, world)
    mine = rays[s:s + c].copy()
    local.intersect(mine, threads=1)
    hits = torch.from_numpy(np.stack([mine["t"].view(np.int32), mine["prim"].view(np.int32)], 1).copy())
    sizes = [multi.shard_range(rays.shape[0], r, world)[1] for r in range(world)]
    bufs = [torch.empty((k, 2), dtype=torch.int32) for k in sizes] if rank == 0 else None
    dist.gather(hits, bufs, 0)
    if rank == 0:
        whole = rays.copy()
        portpy.PortBVH(v).intersect(whole, threads=1)
        cat = torch.cat(bufs).numpy()
        assert np.array_equal(cat[:, 0], whole["t"].view(np.int32)) and np.array_equal(cat[:, 1], whole["prim"].view(np.int32))
        print("MULTI_OK", world, rays.shape[0])
    dist.barrier()
    dist.destroy_process_group()
""")


def test_world_size_2_gloo_broadcast_and_shard(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, TBVH_REPO=REPO, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_OK 2" in r.stdout


def test_block_cyclic_partition_covers_the_set_once():
    """bench.py --gpus N deals the one ray set out in 2^20-ray blocks, block b to rank b % N."""
    from tinybvh_b200 import multi
    for n in (0, 5, 1 << 20, (1 << 20) + 1, 67108864, 67108864 + 12345):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                blocks = multi.block_cyclic(n, r, world)
                assert all(c > 0 and a % 32 == 0 for a, c in blocks)
                seen += blocks
            seen.sort()
            pos = 0
            for a, c in seen:
                assert a == pos
                pos += c
            assert pos == n
            if world > 1 and n >= world << 20:
                sizes = [sum(c for _, c in multi.block_cyclic(n, r, world)) for r in range(world)]
                assert max(sizes) - min(sizes) <= 1 << 20
