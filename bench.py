#!/usr/bin/env python
"""bench.py - the hot path on N B200s of one node: BVH build once, then per step one pass of primary closest-hit rays
plus one pass of shadow any-hit rays over the resident BVH (BASELINE.json configs[1]: Crytek Sponza, 16M primary rays;
the metric is "Mrays/s (primary+shadow)").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scene sponza] [--layout bvh|cwbvh]

Prints ONE JSON line (rank 0).  `value` = rays of all ranks / max-over-ranks device time with rays resident in HBM;
`e2e` = the same passes through the C-ABI host-buffer calls (pinned 128-byte host Ray records in, hits / occlusion
bits back to host inside the timed region); `roofline` = the closest-hit kernel against measured HBM copy bandwidth;
`cpu_baseline` = the reference's BVH8_CPU (AVX2) on the host cores (oracle/_ref), the oracle port if that is absent.
--impl reference times that CPU path alone (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from tinybvh_b200 import rays as R, scenes  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------- workload
def camera_for(scene, verts, k):
    if scene == "sponza":
        return R.SPONZA_EYES[k % 3], R.SPONZA_VIEWS[k % 3]
    lo, hi = scenes.scene_bounds(verts)
    return R.bounds_camera(lo, hi, "inside" if scene == "bistro" else "outside")


def light_for(scene, verts):
    if scene == "sponza":
        return np.zeros(3, np.float32)  # tiny_bvh_speedtest.cpp:856
    lo, hi = scenes.scene_bounds(verts)
    return ((lo + hi) * 0.5 + np.array([0, (hi - lo)[1] * 0.45, 0], np.float32)).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference(scene, verts, prim, shadow, passes, threads=0, hq=True):
    """The reference's own CPU implementation of the path on the host cores: BVH8_CPU::Build + Intersect / IsOccluded
    (tiny_bvh.h:7210-7472) from oracle/_ref; the pinned plain-C port of BVH::Intersect when _ref is absent."""
    from oracle import portpy, refpy
    n = prim.shape[0]
    if refpy.available():
        kind, cores = "reference", refpy.hardware_threads() if threads <= 0 else threads
        t0 = time.perf_counter()
        bvh = refpy.RefBVH8CPU(verts, hq=hq)
        build_s = time.perf_counter() - t0
        impl = f"BVH8_CPU::{'BuildHQ' if hq else 'Build'} + Intersect/IsOccluded (AVX2), 10k-ray batches off an atomic counter"
    else:
        kind, cores = "port", os.cpu_count() if threads <= 0 else threads
        t0 = time.perf_counter()
        bvh = portpy.PortBVH(verts)
        build_s = time.perf_counter() - t0
        impl = "oracle port of BVH::Build + BVH::Intersect/IsOccluded (scalar C, pthreads)"
    a, s = prim.copy(), shadow.copy()
    bvh.intersect(a, threads), bvh.occluded(s, threads)  # warm-up pass (tiny_bvh_speedtest.cpp:185-215)
    times = []
    for _ in range(passes):
        R.reset_hits(a)
        t0 = time.perf_counter()
        bvh.intersect(a, threads)
        t1 = time.perf_counter()
        bvh.occluded(s, threads)
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
    tp, ts = float(np.mean([t[0] for t in times])), float(np.mean([t[1] for t in times]))
    return {"value": 2 * n / (tp + ts) / 1e6, "unit": "Mrays/s", "cores": int(cores), "kind": kind,
            "sample": f"{n} primary + {n} shadow rays of the workload, {passes} timed passes after 1 warm-up, all host threads",
            "impl": impl, "primary_mrays": n / tp / 1e6, "shadow_mrays": n / ts / 1e6, "build_s": build_s,
            "build_mtris": verts.shape[0] / 3 / build_s / 1e6, "ms_per_step": (tp + ts) * 1e3}


def host_primary_and_shadow(scene, verts, label, res, cam_index):
    """Ray sets on the host: primary rays of one camera, traced once by the CPU reference to derive the shadow rays
    (tiny_bvh_speedtest.cpp:844-865).  Used by the reference arm only (our arm traces on the GPU)."""
    from oracle import portpy, refpy
    eye, view = camera_for(scene, verts, cam_index)
    prim = R.primary_rays(eye, view, res, res, 16)
    o = refpy.RefBVH(verts, mode=0, threaded=True) if refpy.available() else portpy.PortBVH(verts)
    traced = prim.copy()
    o.intersect(traced, 0)
    lo, hi = scenes.scene_bounds(verts)
    sh = R.shadow_rays(traced, light_for(scene, verts), float((hi - lo).max() * 5e-7))
    return prim, sh


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    verts, label = scenes.load_scene(args.scene)
    res = args.res
    # bounded sample: the CPU gets a quarter-resolution slice of the workload per step unless --full-reference
    sres = res if args.full_reference else max(256, res // 2)
    prim, sh = host_primary_and_shadow(args.scene, verts, label, sres, 0)
    cb = cpu_reference(args.scene, verts, prim, sh, passes=max(1, args.steps), hq=args.tree == "hq")
    out = {"metric": METRIC, "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": data_label(label), "impl": "reference",
           "config": {"workload": workload_name(args, label), "rays_per_step": 2 * prim.shape[0], "scene_tris": verts.shape[0] // 3,
                      "note": "CPU reference on host cores; each step is a bounded sample of the workload (see cpu_baseline.sample)"},
           "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "impl", "primary_mrays", "shadow_mrays", "build_mtris")},
           "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


def data_label(label):
    if label.startswith("synthetic"):
        return "synthetic procedural scene of the same triangle count (fixture not found)"
    if label == "lucy_dragon_x29":
        return "reference fixtures testdata/lucy.bin + xyzrgb_dragon.bin replicated 29x on a grid (10,145,708 triangles), rays generated synthetically"
    files = "+".join(scenes.SCENES[label][0])
    return f"reference fixture testdata/{files} (triangle soup), rays generated synthetically (speedtest camera pattern)"


def workload_name(args, label):
    return f"{label}_{'sbvh' if args.tree == 'hq' else 'sah'}_{args.layout}_{args.res}x{args.res}x16_primary+shadow"


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tinybvh_b200 import api, _lib
    import ctypes as C

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    verts, label = scenes.load_scene(args.scene)
    ntris = verts.shape[0] // 3

    # ---- build on rank 0 (timed separately: build Mtris/s), then ONE broadcast of the BVH over NVLink
    bvh = api.BVH(device=local)
    build_ms, bcast_ms, build_hq = None, None, None
    if rank == 0:
        # both builders of the path are timed (second call each: the first pays allocations and first-launch costs); the rays are
        # traced through the tree --tree names - the SBVH by default, as in the reference's own traversal benchmarks
        sah = api.BVH(device=local)
        sah.Build(verts)
        sah = api.BVH(device=local)
        sah.Build(verts)
        build_ms = sah.info().build_ms
        hq = api.BVH(device=local)
        hq.BuildHQ(verts)
        hq = api.BVH(device=local)
        hq.BuildHQ(verts)
        build_hq = {"ms": hq.info().build_ms, "mtris_per_s": ntris / hq.info().build_ms / 1e3, "nodes": hq.info().used_nodes, "idx_count": hq.info().idx_count}
        bvh = hq if args.tree == "hq" else sah
        del hq, sah
    if world > 1:
        from tinybvh_b200 import multi
        arrays = None
        if rank == 0:
            i = bvh.info()
            d_nodes = torch.empty(i.used_nodes * 8, dtype=torch.int32, device=dev)
            d_idx = torch.empty(i.idx_count, dtype=torch.int32, device=dev)
            api.check(L.tbvh_download_bvh(bvh.h, C.c_void_p(d_nodes.data_ptr()), C.c_void_p(d_idx.data_ptr()), api.DEVICE))
            arrays = {"nodes": d_nodes, "prim_idx": d_idx, "verts": torch.from_numpy(verts.reshape(-1)).to(dev)}
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        got = multi.broadcast_arrays(arrays, 0, dev)   # the ONE exchange step: BVH replica to every GPU over NVLink
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            api.check(L.tbvh_upload_bvh(bvh.h, C.c_void_p(got["nodes"].data_ptr()), got["nodes"].numel() // 8, C.c_void_p(got["prim_idx"].data_ptr()),
                                        got["prim_idx"].numel(), C.c_void_p(got["verts"].data_ptr()), 16, ntris, api.DEVICE))
    eng = bvh
    if args.layout == "cwbvh":
        api.check(L.tbvh_convert(bvh.h, api.LAYOUT_CWBVH))
        eng.layout = api.LAYOUT_CWBVH  # same handle, traverse its CWBVH layout
    info = bvh.info()

    # ---- this rank's shard of rays (weak scaling: every rank traces res*res*16 primary + as many shadow rays)
    # every rank traces the same view: per-GPU work is identical, so N-GPU numbers measure the system, not the camera
    eye, view = camera_for(args.scene, verts, 0)
    t0 = time.time()
    prim = R.primary_rays(eye, view, args.res, args.res, 16)
    n = prim.shape[0]
    h_prim = api.pinned_empty(n, R.RAY_DTYPE)
    h_prim[:] = prim
    del prim
    d_prim = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_prim.copy_(torch.from_numpy(h_prim.view(np.uint8).reshape(n, 128)[:, :64]))
    d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    eng.Intersect(d_prim, hits=d_hits)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy()
    traced = h_prim.copy()
    traced["t"], traced["u"], traced["v"], traced["prim"] = hits[:, 0], hits[:, 1], hits[:, 2], hits[:, 3].view(np.uint32)
    lo, hi = scenes.scene_bounds(verts)
    h_shadow = api.pinned_empty(n, R.RAY_DTYPE)
    h_shadow[:] = R.shadow_rays(traced, light_for(args.scene, verts), float((hi - lo).max() * 5e-7))
    del traced
    d_shadow = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_shadow.copy_(torch.from_numpy(h_shadow.view(np.uint8).reshape(n, 128)[:, :64]))
    d_bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
    h_bits = np.zeros((n + 31) // 32, np.uint32)
    if rank == 0:
        log(f"[bench] {label}: {ntris} tris, {info.used_nodes} nodes, depth {info.max_depth}, build {build_ms} ms; {n} primary + {n} shadow rays/rank; setup {time.time() - t0:.1f}s")

    stream = torch.cuda.current_stream(dev)

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        eng.Intersect(d_prim, hits=d_hits)
        if ev is not None:
            ev[1].record(stream)
        eng.IsOccluded(d_shadow, bits=d_bits)
        if ev is not None:
            ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = api.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(args.steps):
        step(evs[k])
    t_end.record(stream)
    barrier()
    launches = api.launch_count() - launches0
    total_ms = t_begin.elapsed_time(t_end)
    prim_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    shad_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))

    # ---- e2e: the reference-facing C-ABI calls on HOST buffers (pinned), copies inside the timed region
    for _ in range(2):
        eng.Intersect(h_prim)
        eng.IsOccluded(h_shadow, bits=h_bits)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        eng.Intersect(h_prim)
        eng.IsOccluded(h_shadow, bits=h_bits)
    te1 = time.perf_counter()
    e2e_ms = (te1 - te0) * 1e3
    # the same with the packed-hits entry point (tbvh_intersect_packed): the return trip is one contiguous copy per chunk
    h_hits = api.pinned_empty(n, R.HIT_DTYPE)
    eng.IntersectPacked(h_prim, hits=h_hits)
    barrier()
    tp0 = time.perf_counter()
    for _ in range(args.steps):
        eng.IntersectPacked(h_prim, hits=h_hits)
        eng.IsOccluded(h_shadow, bits=h_bits)
    tp1 = time.perf_counter()
    e2e_packed_ms = (tp1 - tp0) * 1e3
    clk = clocks.stop() if clocks else None

    t = torch.tensor([total_ms, e2e_ms, prim_ms, shad_ms, e2e_packed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, prim_ms, shad_ms, e2e_packed_ms = [float(x) for x in t.cpu()]

    if rank == 0:
        rays_per_step = 2 * n * world
        ms_per_step = total_ms / args.steps
        value = rays_per_step / ms_per_step / 1e3
        peak, which = peaks()
        if args.layout == "cwbvh":
            bvh_bytes = info.used_blocks * 16 + info.cwbvh_tri_count * 48
        else:
            bvh_bytes = info.used_nodes * 32 + info.idx_count * 48
        alg_bytes = n * 80 + bvh_bytes  # 64 B ray read + 16 B hit write per ray + one pass over the BVH
        achieved = alg_bytes / (prim_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.isfile(tp):
            try:
                traffic = json.load(open(tp)).get(f"k_trace_{'cwbvh' if args.layout == 'cwbvh' else 'bvh2'}_closest_{label}" + ("_sbvh" if args.tree == "hq" else ""))
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": data_label(label),
            "config": {"workload": workload_name(args, label), "scene_tris": ntris, "layout": args.layout,
                       "rays_per_step_per_gpu": 2 * n, "primary_rays_per_gpu": n, "shadow_rays_per_gpu": n,
                       "parallelism": f"rays sharded by index over {world} GPU(s) (each shard = the same 16.8M-ray view), BVH built on rank 0 and broadcast once (NCCL), no collective during traversal" if world > 1 else "1 GPU",
                       "l2": "no flush: per-step inputs (2 x %.2f GB ray records) exceed the 126 MB L2" % (n * 64 / 1e9),
                       "tree": "BVH::BuildHQ (SBVH), as tiny_bvh_speedtest.cpp builds for its traversal runs" if args.tree == "hq" else "BVH::Build (binned SAH)",
                       "bvh_built_on": "GPU (tbvh_build_flavour)"},
            "primary_mrays": n * world / prim_ms / 1e3, "shadow_mrays": n * world / shad_ms / 1e3,
            "build": {"ms": build_ms, "mtris_per_s": (ntris / build_ms / 1e3) if build_ms else None, "bcast_ms": bcast_ms,
                      "bvh_bytes": bvh_bytes, "build_hq": build_hq},
            "roofline": {"bound": "hbm", "kernel": "k_trace_cwbvh<closest>" if args.layout == "cwbvh" else "k_trace_bvh2<closest>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": which, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": prim_ms,
                         "note": "traversal is latency/issue bound, not HBM bound (SURVEY 8d): see DESIGN.md for the L2-side accounting"},
            "e2e": {"value": rays_per_step / (e2e_ms / args.steps) / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n * 64,
                    "d2h_bytes_per_step": n * 16 + ((n + 31) // 32) * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "tbvh_intersect + tbvh_occluded on pinned 128-byte host Ray records (hits written in place into Ray.hit)",
                    "packed_hits_value": rays_per_step / (e2e_packed_ms / args.steps) / 1e3,
                    "packed_hits_api": "tbvh_intersect_packed + tbvh_occluded: same inputs, hits returned as a packed 16-byte array"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.time()
            sres = max(256, args.res // 2)
            p2, s2 = host_primary_and_shadow(args.scene, verts, label, sres, 0)
            cb = cpu_reference(args.scene, verts, p2, s2, passes=3, hq=args.tree == "hq")
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "impl", "primary_mrays", "shadow_mrays", "build_mtris")}
            log(f"[bench] cpu baseline took {time.time() - t0:.1f}s")
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="sponza")
    ap.add_argument("--layout", default="bvh", choices=["bvh", "cwbvh"])
    ap.add_argument("--tree", default="hq", choices=["hq", "sah"],
                    help="hq: BVH::BuildHQ (SBVH), the tree every traversal benchmark of tiny_bvh_speedtest.cpp builds (:894, 960, 1013, 1099, 1197); sah: BVH::Build")
    ap.add_argument("--res", type=int, default=1024, help="primary rays = res*res*16 (1024 -> 16,777,216)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-reference", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
