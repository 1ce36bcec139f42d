#!/usr/bin/env python3
"""Key counters of every kernel in an ncu report (--set full): duration, instructions, issue utilisation, L1 data-pipe
wavefronts by source, pipe utilisation, stall reasons, DRAM bytes, L2 atomics.
  python tools/ncu_key_metrics.py <report.ncu-rep>"""
import csv, io, re, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
pat = re.compile(r"^(gpu__time_duration.sum|launch__registers_per_thread|launch__occupancy_limit_(registers|shared_mem)|sm__warps_active.avg.pct_of_peak_sustained_active|"
                 r"smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed.ratio|"
                 r"l1tex__data_pipe_lsu_wavefronts.sum|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|"
                 r"l1tex__data_pipe_lsu_wavefronts_mem_shared_op_(ld|st).sum|l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|red|st).sum|l1tex__t_requests_pipe_lsu_mem_global_op_(ld|red|st).sum|"
                 r"l1tex__t_requests_pipe_lsu_mem_local_op_(ld|st).sum|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|lts__t_sectors_op_(red|atom).sum|"
                 r"lts__t_sectors_srcunit_tex_op_red.sum|lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed|lts__t_sectors_srcunit_tex_op_red.sum.per_second|lts__t_sectors_srcunit_tex_op_red_lookup_(hit|miss).sum|lts__t_sectors_srcunit_tex_op_atom.sum|lts__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"sm__inst_executed_pipe_(alu|fma|lsu|xu|cbu|uniform).avg.pct_of_peak_sustained_active|dram__bytes_(read|write).sum|sm__cycles_elapsed.max|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__warps_eligible.avg.per_cycle_active)$")
keys = [i for i, h in enumerate(hdr) if pat.match(h)]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    stalls = []
    for i in keys:
        h = hdr[i]
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m:
            try: stalls.append((float(r[i]), m.group(1)))
            except ValueError: pass
            continue
        print(f"   {h:78s} {r[i]} {rows[1][i]}")
    print("   stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
