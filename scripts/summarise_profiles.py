"""Turns the ncu outputs of scripts/gpu_profile_round.sh into the committed summaries:

    profiles/<round>_launches_<cfg>.md     per-kernel share of a bench step (from the launch list)
    profiles/<round>_ncu_<kernel>.txt      key metrics + hottest SASS lines of the full capture
    profiles/ncu_traffic.json              dram bytes per launch, read by bench.py (`roofline.traffic`)
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = ROOT / "profiles"
OUT.mkdir(exist_ok=True)


def launches(cfg):
    p = ROOT / "gpurun_out" / f"launches_{R}_{cfg}.csv"
    if not p.is_file():
        return
    rows = list(csv.reader(l for l in p.read_text().splitlines() if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        tot[name] += float(r[iv].replace(",", "")) / 1e3     # ns -> us
        cnt[name] += 1
    ours = {k: v for k, v in tot.items() if "srcv" in k or "unnamed" in k}
    s_ours = sum(ours.values())
    lines = [f"# {R} launch list, bench.py workload {cfg} (ncu --metrics gpu__time_duration.sum, cold & serialised:",
             "# compare SHARES, not absolutes)", "", "| kernel | launches | total us | share of our kernels |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        share = f"{100 * v / s_ours:.1f} %" if k in ours else "(torch)"
        lines.append(f"| `{k[-70:]}` | {cnt[k]} | {v:.1f} | {share} |")
    (OUT / f"{R}_launches_{cfg}.md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(tag, traffic_key, traffic):
    rep = ROOT / "gpurun_out" / f"prof_{R}_{tag}.ncu-rep"
    if not rep.is_file():
        return
    a = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_hot.py"), str(rep), "25"], capture_output=True, text=True).stdout
    (OUT / f"{R}_ncu_{tag}.txt").write_text(f"# ncu --set full --clock-control none, {rep.name}\n" + a + "\n" + b)
    print(a)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic[traffic_key] = int(get("dram__bytes_read.sum") + get("dram__bytes_write.sum"))


launches("cfg1")
launches("cfg2")
tp = OUT / "ncu_traffic.json"
traffic = json.loads(tp.read_text()) if tp.is_file() else {}
full("dot", "dot:dot_fast_c4planar", traffic)
full("hero", "mlp:mlp_tc_tcgen05_f16x3", traffic)
tp.write_text(json.dumps(traffic, indent=1))
print(traffic)
