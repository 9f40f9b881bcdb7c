"""Summarise an .ncu-rep (read on the CPU box): per-kernel duration, DRAM bytes, utilisation, stalls.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]])
    for w in want:
        if w in idx:
            print(f"  {w:75s} {r[idx[w]]:>18s} {units[idx[w]]}")
    top = sorted(((float(r[idx[s]].replace(',', '') or 0), s) for s in stalls), reverse=True)[:6]
    for v, s in top:
        print(f"  stall {s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f} warps/issue")
