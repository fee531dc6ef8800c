"""Per-source-line view of an ncu report (needs --import-source on and -lineinfo): executed warp instructions and
stall samples aggregated by CUDA source line.  usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None; hdr = None
tot_i = tot_s = 0
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and len(r) > iI and r[0].isdigit():
        i, s = int(r[iI]) if r[iI].isdigit() else 0, int(r[iS]) if r[iS].isdigit() else 0
        lines.append((cur, int(r[0]), r[1].strip(), i, s)); tot_i += i; tot_s += s
print(f"total warp instructions {tot_i}  samples {tot_s}")
lines.sort(key=lambda x: -x[3])
for f, ln, src, i, s in lines[:top]:
    print(f"{f}:{ln:<4d} {100*i/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% samp  {src[:100]}")
