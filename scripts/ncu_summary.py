"""Selected raw metrics of an ncu report as 'name,value,unit' lines (what profiles/*_ncu.csv hold).
usage: python scripts/ncu_summary.py report.ncu-rep "title line" > profiles/xxx_ncu.csv"""
import csv, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration", "dram__bytes", "dram__throughput", "launch__", "smsp__issue_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_", "sm__warps_active", "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate",
        "l1tex__data_bank_conflicts", "sm__throughput", "smsp__warps_eligible", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed")
print(f"# {title}")
print("# ncu --set full --clock-control none (raw page, selected metrics)")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keep) and not (".min" in h or ".max.pct" in h):
        print(f"{h},{v},{u}")
