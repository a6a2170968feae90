"""Static SASS instruction count per CUDA source line of one kernel (needs -lineinfo).

    python scripts/sass_lines.py <file.o|.so|.cubin> <kernel-name-substring> [source.cu] [N]

Complements scripts/ncu_lines.py (dynamic counts from an ncu report): this one needs no GPU.
Lines inside loops count once; read it next to the loop structure.
"""
import re
import subprocess
import sys
import tempfile
from collections import Counter
from pathlib import Path

obj, pat = sys.argv[1], sys.argv[2]
srcfile = sys.argv[3] if len(sys.argv) > 3 else None
n = int(sys.argv[4]) if len(sys.argv) > 4 else 50
with tempfile.TemporaryDirectory() as td:
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, check=True,
                       capture_output=True)
        cubins = sorted(Path(td).glob("*.cubin"))
    else:
        cubins = [Path(obj)]
    txt = ""
    for c in cubins:
        txt += subprocess.run(["nvdisasm", "--print-line-info", str(c)], capture_output=True, text=True).stdout
sections = re.split(r"\n//-+ \.text\.", txt)
src = open(srcfile).read().split("\n") if srcfile else None
for sec in sections[1:]:
    name = sec.split(" ", 1)[0]
    if pat not in name:
        continue
    cur, cnt, inl = None, Counter(), Counter()
    for l in sec.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l) and cur:
            cnt[cur] += 1
    tot = sum(cnt.values())
    print(f"== {name[:100]}: {tot} instructions")
    for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:n]:
        t = src[ln - 1].strip()[:100] if src and srcfile.endswith(f) else ""
        print(f"{c:5d} {f}:{ln} {t}")
