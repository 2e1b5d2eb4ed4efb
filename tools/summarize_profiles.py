#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun session left in gpurun_out/ into the small tracked summaries under profiles/.
usage: tools/summarize_profiles.py <round-tag> [launches.csv] [trace.ncu-rep] [build.ncu-rep]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
           "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active"]


def launches(tag, path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in r:
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        v = float(row[vi].replace(",", ""))
        tot[name] = tot.get(name, 0) + v
        cnt[name] += 1
    T = sum(tot.values())
    with open(os.path.join(OUT, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(cnt.values())} launches, {T / 1e6:.3f} ms of kernel time (cold-cache, serialised)\n")
        f.write(f"# {'kernel':58s} {'n':>5s} {'total ms':>10s} {'share %':>8s} {'avg us':>10s}\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{k:60s} {cnt[k]:5d} {v / 1e6:10.3f} {100 * v / T:8.2f} {v / cnt[k] / 1e3:10.1f}\n")
    subprocess.call(["cp", path, os.path.join(OUT, f"{tag}_launches.csv")])


def report(tag, name, path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {}
    with open(os.path.join(OUT, f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# extract of `ncu --set full --clock-control none` ({os.path.basename(path)}); one block per captured launch\n")
        for r in rows[2:]:
            kn = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")
            f.write(f"\n== {kn}  grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''} block {r[hdr.index('Block Size')] if 'Block Size' in hdr else ''}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:75s} {r[i]:>18s} {units[i]}\n")
            if "dram__bytes_read.sum" in hdr:
                def tobytes(i):
                    v, u = float(r[i].replace(",", "")), units[i]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                out.setdefault(kn, []).append(tobytes(hdr.index("dram__bytes_read.sum")) + tobytes(hdr.index("dram__bytes_write.sum")))
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and os.path.isfile(sys.argv[2]):
        launches(tag, sys.argv[2])
    traffic = {}
    for name, arg in (("trace", 3), ("build", 4)):
        if len(sys.argv) > arg and os.path.isfile(sys.argv[arg]):
            t = report(tag, name, sys.argv[arg])
            for k, v in t.items():
                traffic[k] = v
    print(json.dumps(traffic, indent=1))
