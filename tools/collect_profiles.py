#!/usr/bin/env python
"""Copy what a tools/gpu_session_full.sh run left in gpurun_out/ into tracked files under profiles/ (tag = round, e.g. r2)."""
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sys.path.insert(0, os.path.join(REPO, "tools"))
import summarize_profiles as sp  # noqa: E402


def line(src, dst):
    p = os.path.join(G, src)
    if not os.path.isfile(p):
        return None
    t = open(p).read()
    i = t.find('{"metric"')
    if i < 0:
        return None
    d = json.loads(t[i:].splitlines()[0])
    json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
    return d


names = {"bench_n1.json": f"{tag}_bench_bistro_cwbvh_n1.json", "bench_ref.json": f"{tag}_bench_reference_arm.json", "bench_n1_bvh.json": f"{tag}_bench_bistro_bvh_n1.json",
         "bench_sponza_bvh.json": f"{tag}_bench_sponza_bvh_n1.json", "bench_sponza_cwbvh.json": f"{tag}_bench_sponza_cwbvh_n1.json", "bench_config5.json": f"{tag}_bench_config5_lucy_dragon_x29_n1.json",
         "bench_n2.json": f"{tag}_bench_bistro_cwbvh_n2.json", "bench_n4.json": f"{tag}_bench_bistro_cwbvh_n4.json", "bench_n8.json": f"{tag}_bench_bistro_cwbvh_n8.json", "bench_config4_n8.json": f"{tag}_bench_config4_bistro_537M_n8.json"}
for src, dst in names.items():
    d = line(src, dst)
    if d:
        print(f"{dst}: value {d['value']:.0f} e2e {d['e2e']['value']:.0f} {d['unit']}")
if os.path.isfile(os.path.join(G, "launches.csv")):
    sp.launches(tag, os.path.join(G, "launches.csv"))
traffic = {}
tp = os.path.join(P, "traffic.json")
if os.path.isfile(tp):
    traffic = json.load(open(tp))
for name, rep in (("cw_primary", f"{tag}_cw_primary.ncu-rep"), ("cw_shadow", f"{tag}_cw_shadow.ncu-rep"), ("large_phase", f"{tag}_large_phase.ncu-rep"),
                  ("cw_primary_2048", f"{tag}_cw_primary_2048.ncu-rep")):
    p = os.path.join(G, rep)
    if os.path.isfile(p):
        t = sp.report(tag, name, p)
        for k, v in t.items():
            traffic[f"{tag}:{name}:{k}"] = v[0]
        if name == "cw_primary_2048":
            for v in t.values():
                traffic["k_trace_wide<closest>_bistro_hq_2048"] = v[0]   # what bench.py's default line looks up
        elif name != "large_phase":
            subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_opmix.py"), p, "16777216", os.path.join(P, f"{tag}_{name}_opmix.txt")], stdout=subprocess.DEVNULL)
# bench.py looks the dominant kernel's DRAM bytes up by "<kernel>_<scene>_<tree>_<res>"; the capture is the 1024^2 x 16 camera set
k = f"{tag}:cw_primary:k_trace_wide<0, 0, 1>"
if k in traffic:
    traffic["k_trace_wide<closest>_bistro_hq_1024"] = traffic[k]
json.dump(traffic, open(tp, "w"), indent=1)
for f in ("build.log", "hq.log"):
    if os.path.isfile(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f"{tag}_{f.replace('.log', '')}_times.txt"))
print("traffic:", json.dumps(traffic, indent=1))
