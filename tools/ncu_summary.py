"""tools/ncu_summary.py REPORT.ncu-rep [kernel-substring]: the counters DESIGN.md / profiles quote, from `ncu --set full` reports
(ncu -i ... --page raw --csv), one block per captured launch."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    if want and want not in r[ik]:
        continue
    print("Kernel =", r[ik][:120])
    vals = dict(zip(hdr, r))
    un = dict(zip(hdr, units))
    for k in KEYS:
        if k in vals and vals[k] != "":
            print(f"  {k} = {vals[k]} {un.get(k, '')}")
    stalls = sorted(((float(vals[h].replace(',', '')), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h in hdr
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals.get(h, "") not in ("", "n/a")), reverse=True)
    print("  stalls (warps per issued instruction) =", ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]))
    print()
