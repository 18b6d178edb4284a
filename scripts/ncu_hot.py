#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, and (with --kernel NAME [--index i]) the SASS
regions grouped by execution count with their instruction and stall-sample shares."""
import argparse, csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout

def raw(rep):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print("----")
        for i in idx:
            print(f"  {hdr[i]} [{units[i]}] = {r[i]}")

def source(rep, kernel, index, thresh):
    txt = run(["-i", rep, "--page", "source", "--csv", "--kernel-name", kernel])
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []; blocks.append(cur); continue
        if cur is not None:
            cur.append(r)
    b = blocks[index]
    hdr = b[0]
    si, ii, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(r[si].strip(), int(r[ii] or 0), int(r[ss] or 0)) for r in b[1:] if len(r) > ii]
    tot, st = sum(d[1] for d in data), sum(d[2] for d in data)
    print("instances", len(blocks), "total inst", tot, "samples", st, "sass", len(data))
    i = 0
    while i < len(data):
        j = i
        while j < len(data) and abs(data[j][1] - data[i][1]) <= 0.03 * max(data[i][1], 1):
            j += 1
        cnt, samp = sum(d[1] for d in data[i:j]), sum(d[2] for d in data[i:j])
        ops = {}
        for d in data[i:j]:
            t = d[0].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:6]
        if cnt / tot > thresh or samp / max(st, 1) > thresh:
            print(f"[{i:4d}-{j-1:4d}] n={j-i:3d} exec~{data[i][1]:9d} inst={100*cnt/tot:5.1f}% samp={100*samp/max(st,1):5.1f}% {top}")
        i = j

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--kernel")
ap.add_argument("--index", type=int, default=0)
ap.add_argument("--thresh", type=float, default=0.015)
a = ap.parse_args()
if a.kernel:
    source(a.rep, a.kernel, a.index, a.thresh)
else:
    raw(a.rep)
