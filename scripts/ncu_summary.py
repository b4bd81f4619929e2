#!/usr/bin/env python
"""Summarise an .ncu-rep of the step kernel: key raw metrics + stall samples per barrier-delimited
phase (from the source page).  usage: scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
        print(f"{h} [{u}] = {v}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = rows[2:]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data) or 1
print(f"\nstall samples per barrier-delimited segment (total {tot}):", file=out)
start = samp = dm = 0
for i, r in enumerate(data):
    samp += int(r[isamp])
    dm += "DMMA" in r[isrc]
    if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc]:
        if samp * 200 > tot:
            print(f"  sass[{start:4d}:{i:4d}] {100.0 * samp / tot:5.1f}%  dmma_instrs={dm:3d}  ends: {r[isrc].strip()[:48]}", file=out)
        start, samp, dm = i + 1, 0, 0
print("\ntop instructions by samples:", file=out)
for r in sorted(data, key=lambda r: -int(r[isamp]))[:25]:
    print(f"  {int(r[isamp]):7d} exec={r[iexec]:>10s}  {r[isrc].strip()[:100]}", file=out)
