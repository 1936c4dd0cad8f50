#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-instruction stall samples) into text.  Usage: ncu_summary.py rep [topN]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90], "block", d.get("Block Size"), "grid", d.get("Grid Size"))
    for w in want:
        if w in d:
            print(f"  {w:86s} {d[w]} {units[hdr.index(w)]}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            if v >= 0.05:
                print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    hdr = rows[hi[0]]
    data = []
    for r in rows[hi[0] + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        data.append(r)
    isamp, iexe, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[isamp]) for r in data) or 1
    texe = sum(int(r[iexe]) for r in data) or 1
    print(f"-- {len(data)} SASS instructions, {tot} samples, {texe} warp-instructions executed")
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
    for i in sorted(idx):
        r = data[i]
        print(f"  [{i:5d}] {100*int(r[isamp])/tot:5.1f}%  exe={r[iexe]:>10s}  {r[isrc].strip()[:80]}")
    step = max(100, len(data) // 25)
    for a in range(0, len(data), step):
        s = sum(int(r[isamp]) for r in data[a:a + step]); e = sum(int(r[iexe]) for r in data[a:a + step])
        print(f"  region {a:5d}-{a+step:5d}: samples {100*s/tot:5.1f}%  executed {100*e/texe:5.1f}%")
