#!/usr/bin/env python3
"""Which L1TEX sub-unit bounds a kernel: print the l1tex__ utilisation breakdown of every launch in an .ncu-rep.
usage: tools/ncu_l1.py rep.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_sectors.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
name_col = idx.get("Kernel Name", 4)
for r in rows[2:]:
    if len(sys.argv) > 2 and sys.argv[2] not in r[name_col]:
        continue
    print("==", r[name_col][:60])
    for k in KEYS:
        if k in idx:
            print("   %-95s %s" % (k, r[idx[k]]))
