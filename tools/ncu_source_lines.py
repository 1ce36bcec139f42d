#!/usr/bin/env python3
"""Per-source-line summary of an ncu report captured with --import-source on (kernels compiled with -lineinfo):
  python tools/ncu_source_lines.py <report.ncu-rep> <kernel substring> [min %]
prints, per (file, line): warp instructions executed, share of the kernel, stall samples and their share,
average active threads and the dominant stall reason -- the table the optimisation notes in profiles/ quote."""
import csv
import io
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, cur_fn, hdr = None, None, None
lines = {}     # (file, line) -> dict
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur_fn is None or want not in cur_fn or r[0] == "":
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"]); thr = int(d["Thread Instructions Executed"])
    except (KeyError, ValueError):
        continue
    stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
    e = lines.setdefault((cur_file, int(r[0])), {"src": r[1].strip(), "inst": 0, "samp": 0, "thr": 0, "stalls": {}})
    e["inst"] += inst; e["samp"] += samp; e["thr"] += thr
    for k, v in stalls.items():
        e["stalls"][k] = e["stalls"].get(k, 0) + v
ti = sum(e["inst"] for e in lines.values()); ts = sum(e["samp"] for e in lines.values())
print(f"# {want}: {ti} warp instructions, {ts} stall samples")
print("file:line | inst % | samples % | avg thr | top stall | source")
for (f, ln), e in sorted(lines.items()):
    pi, ps = 100.0 * e["inst"] / max(ti, 1), 100.0 * e["samp"] / max(ts, 1)
    if pi < min_pct and ps < min_pct:
        continue
    top = max(e["stalls"].items(), key=lambda kv: kv[1])[0][6:] if e["stalls"] else "-"
    print(f"{f}:{ln} | {pi:5.2f} | {ps:5.2f} | {e['thr'] / max(e['inst'], 1):4.1f} | {top} | {e['src'][:110]}")
