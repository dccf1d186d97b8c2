"""Summarise an .ncu-rep: key metrics + hot SASS segments.  python tools/ncu_summary.py rep [units_per_warp]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_uniform.avg.pct", "sm__inst_executed_pipe_adu"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want) and "Not Issued" not in h and v not in ("", "0"):
            print(f"  {h:85s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
tot = sum(int(r[5]) for r in rows); samp = sum(int(r[4]) for r in rows)
print("total warp inst", tot, "samples", samp, "sass lines", len(rows))
if len(sys.argv) > 2 and sys.argv[2] == "full":
    for i, r in enumerate(rows):
        print(f"{i:4d} {int(r[5])/1e3:9.0f}k {int(r[4]):6d}  {r[1].strip()[:100]}")
else:
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][4]))[:40]
    print("top stall-sample instructions:")
    for i in sorted(top):
        r = rows[i]
        print(f"{i:4d} inst {int(r[5])/1e3:8.0f}k samples {int(r[4]):6d} ({100*int(r[4])/samp:4.1f}%)  {r[1].strip()[:90]}")
