#!/usr/bin/env python3
"""Attribute the warp instructions of a kernel to REGIONS of its source: SASS rows of an ncu report (--import-source on,
-lineinfo) are sorted by address; rows that belong to inlined helpers (other files, or helper lines of the same file)
inherit the last line of the kernel body seen before them, then lines are bucketed into the given ranges.
  python tools/ncu_regions.py <report> <kernel substring> <file> name:lo-hi [name:lo-hi ...]"""
import csv, io, subprocess, sys
rep, want, body = sys.argv[1:4]
regions = []
for spec in sys.argv[4:]:
    name, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = cur_fn = hdr = None
cur_line = None
sass = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur_fn is None or want not in cur_fn:
        continue
    if r[0] != "":
        cur_line = int(r[0]); continue
    if r[2] in ("...", "-", ""):
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        sass.append((int(r[2], 16), cur_file, cur_line, int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
    except (ValueError, KeyError):
        pass
sass.sort()
lo_all = min(lo for _, lo, _ in regions); hi_all = max(hi for _, _, hi in regions)
last = None
agg = {}
for addr, f, ln, inst, thr, samp in sass:
    if f == body and lo_all <= ln <= hi_all:
        last = ln
    key = "other"
    if last is not None:
        for name, lo, hi in regions:
            if lo <= last <= hi:
                key = name; break
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += inst; a[1] += thr; a[2] += samp
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"# {want}: {ti} warp instructions, {ts} samples")
for name, (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{name:28s} inst {100 * i / ti:5.1f}%  ({i / 1e6:7.2f} M)  samples {100 * s / max(ts, 1):5.1f}%  avg threads {t / max(i, 1):4.1f}")
