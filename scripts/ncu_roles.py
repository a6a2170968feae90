"""Stall samples of a warp-specialised kernel split by ROLE: the SASS regions between the
USETMAXREG instructions (setmaxnreg) that open each role's branch.

    python scripts/ncu_roles.py gpurun_out/prof.ncu-rep
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], {"name": "prologue", "rows": []}
for r in body:
    src = r[ix["Source"]]
    if "USETMAXREG" in src:
        regions.append(cur)
        cur = {"name": src.strip()[:40], "rows": []}
    cur["rows"].append(r)
regions.append(cur)
tot = sum(int(r[ix["# Samples"]]) for r in body)
print(f"total samples {tot}")
for g in regions:
    s = sum(int(r[ix["# Samples"]]) for r in g["rows"])
    inst = sum(int(r[ix["Instructions Executed"]]) for r in g["rows"])
    waits = sum(int(r[ix["# Samples"]]) for r in g["rows"]
                if "TRYWAIT" in r[ix["Source"]] or ("BRA" in r[ix["Source"]] and int(r[ix["Instructions Executed"]]) > 0
                                                    and False))
    # samples on the branch right after a TRYWAIT count as waiting too
    w2 = 0
    rws = g["rows"]
    for i, r in enumerate(rws):
        if "TRYWAIT" in r[ix["Source"]]:
            w2 += int(r[ix["# Samples"]])
            if i + 1 < len(rws) and "BRA" in rws[i + 1][ix["Source"]]:
                w2 += int(rws[i + 1][ix["# Samples"]])
    agg = {h: sum(int(r[ix[h]]) for r in g["rows"]) for h in stall_cols}
    top = ", ".join(f"{h[6:]}={100 * v / max(s, 1):.0f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:5])
    print(f"{g['name']:42s} samples {100 * s / tot:5.1f}%  warp-instr {inst:>11,d}  mbarrier-wait {100 * w2 / max(s, 1):5.1f}% of role | {top}")
