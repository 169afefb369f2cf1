#!/usr/bin/env python
"""Runs HERE (no GPU): turns what scripts/profile_r02.sh brought back in gpurun_out/r02/ into the tracked summaries
under profiles/ (r02_*.json, traffic.json with the commit the capture belongs to)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r02")
DST = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def wide(path):
    """`ncu --page raw --csv`: one row per launch, one column per metric (second row = units)."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                d[h] = {"value": num(v), "unit": u}
        out.append(d)
    return out


def long(path):
    """`ncu --metrics ... --csv`: one row per (launch, metric)."""
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ii], {"kernel": r[ki]})[r[mi]] = {"value": num(r[vi]), "unit": r[ui]}
    return list(d.values())


def derive(k):
    t = k.get("gpu__time_duration.sum", {}).get("value")
    rd = k.get("dram__bytes_read.sum", {}).get("value")
    wr = k.get("dram__bytes_write.sum", {}).get("value")
    scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "msecond": 1e-3, "ms": 1e-3, "nsecond": 1e-9}
    bscale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if isinstance(t, float) and isinstance(rd, float) and isinstance(wr, float):
        ts = t * scale.get(k["gpu__time_duration.sum"]["unit"], 1e-9)
        b = rd * bscale.get(k["dram__bytes_read.sum"]["unit"], 1) + wr * bscale.get(k["dram__bytes_write.sum"]["unit"], 1)
        k["derived"] = {"us": round(ts * 1e6, 2), "dram_MB": round(b / 1e6, 2), "dram_GBps": round(b / ts / 1e9, 1),
                        "frac_of_measured_copy_peak_6554.9": round(b / ts / 1e9 / 6554.9, 4)}
    return k


commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
meta = {"commit": commit, "recipe": "scripts/profile_r02.sh (ncu --cache-control none --clock-control none; PDL off for the --set full "
                                    "captures so that every kernel is measured alone; cold-cache, serialised launches: compare shares "
                                    "and traffic, absolute times come from CUDA events in bench.py)"}
full = [derive(k) for k in wide(os.path.join(SRC, "q1_solver_kernels.raw.csv"))]
json.dump({"meta": meta, "workload": "Q1 (1 003 002 dof), classic PCG, steady state", "kernels": full},
          open(os.path.join(DST, f"{TAG}_ncu_full_q1_kernels.json"), "w"), indent=1)
asm = [derive(k) for k in wide(os.path.join(SRC, "assemble_node_q1.raw.csv"))]
json.dump({"meta": meta, "workload": "Q1 assembly (500 000 quads), node-parallel GATHER", "kernels": asm},
          open(os.path.join(DST, f"{TAG}_ncu_full_assemble_node_q1.json"), "w"), indent=1)
other = {}
for name, f in (("k1_q16", "k1_q16.csv"), ("spmv_l64", "spmv_l64.csv"), ("fem_kernels_q1", "fem_kernels.csv"),
                ("row_partitioned_kernels_world1_q1", "dist_kernels_world1_q1.csv")):
    other[name] = [derive(k) for k in long(os.path.join(SRC, f))]
json.dump({"meta": meta, "captures": other}, open(os.path.join(DST, f"{TAG}_ncu_metrics.json"), "w"), indent=1)
k1 = next(k for k in full if "spmv_stream" in k["kernel"])
rd, wr = k1["dram__bytes_read.sum"], k1["dram__bytes_write.sum"]
bscale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rdb, wrb = int(rd["value"] * bscale[rd["unit"]]), int(wr["value"] * bscale[wr["unit"]])
json.dump({"source": f"ncu --set full --cache-control none --clock-control none, steady state of the Q1 solve "
                     f"(profiles/{TAG}_ncu_full_q1_kernels.json, scripts/profile_r02.sh); re-run after every change of K1",
           "commit": commit, "capture": f"profiles/{TAG}_ncu_full_q1_kernels.json",
           "krylov_spmv_kernel_dram_bytes_per_launch": rdb + wrb, "krylov_spmv_kernel_dram_read_bytes": rdb,
           "krylov_spmv_kernel_dram_write_bytes": wrb,
           "note": "less than the 236.3 MB CSR-model bytes: blocked 16-bit column ids (8.5 instead of 12 B per entry) and work "
                   "vectors resident in L2"}, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
# launch list + shares
src = os.path.join(SRC, "launches_bench_q1.csv")
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    tot[name] += num(r[vi]); cnt[name] += 1
open(os.path.join(DST, f"{TAG}_launches_bench_q1.csv"), "w").write(open(src).read())
total = sum(tot.values())
json.dump({"meta": meta, "launches": int(sum(cnt.values())),
           "shares": {k: {"launches": cnt[k], "ns_total": tot[k], "ns_avg": round(tot[k] / cnt[k], 1), "share": round(tot[k] / total, 4)}
                      for k in tot}}, open(os.path.join(DST, f"{TAG}_launch_shares_bench_q1.json"), "w"), indent=1)
for k in full + asm:
    print(k["kernel"][:70], k.get("derived"))
for name, ks in other.items():
    for k in ks:
        print(name, k["kernel"][:50], k.get("derived"))
