"""Top stalled SASS instructions of an ncu report (needs -lineinfo / --import-source).

    python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in body) for h in stall_cols}
print("total samples", tot)
print("by reason:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("executed warp-instructions:", sum(int(r[ix['Instructions Executed']]) for r in body))
ranked = sorted(enumerate(body), key=lambda ir: -int(ir[1][ix["# Samples"]]))[:n]
for i, r in sorted(ranked):
    s = int(r[ix["# Samples"]])
    top = max(stall_cols, key=lambda h: int(r[ix[h]]))
    print(f"{i:5d} {100*s/tot:5.1f}%  {top[6:]:12s} ex={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()}")
