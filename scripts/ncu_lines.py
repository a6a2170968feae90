"""Executed warp-instructions and stall samples per CUDA source line of an ncu report
(needs -lineinfo + --import-source on).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, ix = None, None, None
lines = []   # (file, line, text, inst, samples)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
    elif hdr and len(r) == len(hdr) and r[0] != "":
        lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ix["Instructions Executed"]] or 0),
                      int(r[ix["# Samples"]] or 0)))
tot_i = sum(l[3] for l in lines)
tot_s = sum(l[4] for l in lines)
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
for f, ln, txt, inst, smp in sorted(lines, key=lambda l: -l[3])[:n]:
    print(f"{100*inst/tot_i:5.1f}%i {100*smp/max(tot_s,1):5.1f}%s  {f}:{ln:<4d} {txt[:110]}")
