#!/usr/bin/env python3
"""Turns the scratch artefacts of a gpurun session (gpurun_out/) into the tracked summaries under profiles/:
  python tools/summarize_profiles.py <tag> <launches.csv> <prof.ncu-rep>
writes profiles/<tag>_launches_bench_summary.md, appends a section to profiles/<round of the tag>_ncu_summaries.md and
refreshes profiles/traffic.json (dram read+write bytes per launch of the two big kernels)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else tag

# ---- launch list ----
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("==")) if r]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1; a[1] += v
ours = {k: v for k, v in agg.items() if "gvv::" in k and "atomic_" not in k}
tot = sum(v[1] for v in ours.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches_bench_summary.md"), "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 4 --warmup 3` (gpu__time_duration.sum, --clock-control none), {title}\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's CUDA-event kernel_ms_per_step\n\n")
    f.write("| kernel | launches | avg us | share of our kernels |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{100 * t / tot:5.1f}%" if k in ours else "-"
        f.write(f"| `{k[:70]}` | {n} | {t / n:.1f} | {share} |\n")

# ---- ncu full-set summary ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_red.sum.per_second", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
traffic = {}
summary_md = os.path.join(ROOT, "profiles", f"{tag[:3]}_ncu_summaries.md")
if not os.path.exists(summary_md):
    open(summary_md, "w").write(f"# ncu --set full summaries, round {int(tag[1:3])} (B200, config 2: 8 views, 1024x1024, 70k tris)\n"
                                "# `ncu --set full --clock-control none --import-source on -k regex:\"raster_kernel|pixel_grad\" --launch-skip 2 --launch-count 2 python tools/gpu_profile_target.py`\n"
                                "# (per-kernel counters; L2 atomics = lts__t_sectors_srcunit_tex_op_red, percentage of the L2's own peak)\n")
with open(summary_md, "a") as f:
    f.write(f"\n## {tag}: {title}\n")
    for r in rr[2:]:
        d = dict(zip(h, r))
        units = dict(zip(h, rr[1]))
        name = d["Kernel Name"]
        f.write(f"\n### {name}\n\n| metric | value |\n|---|---|\n")
        for w in want:
            if w in d and d[w] != "":
                f.write(f"| {w} | {d[w]} {units.get(w, '')} |\n")
        st = sorted(((float(d[k] or 0), k) for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)[:8]
        f.write("| top stall reasons (warps per issue) | " + ", ".join(f"{k.split('stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in st) + " |\n")
        def byts(key):
            v = float(d[key].replace(",", "")); u = units[key]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        short = "raster_kernel" if "raster_kernel" in name else ("pixel_grad_kernel" if "pixel_grad" in name else None)
        if short:
            traffic[short] = int(byts("dram__bytes_read.sum") + byts("dram__bytes_write.sum"))
            traffic[short + "_warp_inst"] = int(float(d["smsp__inst_executed.sum"].replace(",", "")))
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("wrote", tag, traffic)
