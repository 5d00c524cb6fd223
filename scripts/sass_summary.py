"""cuobjdump -sass of libeemflow_b200.so -> per-kernel counts of the Blackwell-specific SASS mnemonics
(profiles/r02/sass_summary.md).  Usage: python scripts/sass_summary.py > /tmp/table.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "eemflow_b200" / "libeemflow_b200.so"
COLS = [("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA .2CTA", r"UTCHMMA\.2CTA"), ("UTMALDG", r"UTMALDG"), ("UTMALDG .MULTICAST", r"UTMALDG\S*MULTICAST"),
        ("LDTM", r"\bLDTM"), ("UTCBAR", r"UTCBAR"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"LDGSTS"), ("STG .256", r"STG\S*\.256"),
        ("F2FP .SATFINITE.F16", r"F2FP\S*SATFINITE\S*F16"), ("MATCH.ANY", r"MATCH\.ANY"), ("ATOMS .CAST.SPIN", r"ATOMS\.CAST\.SPIN"),
        ("ATOMS.MIN", r"ATOMS\.MIN"), ("RED(G).ADD.F32", r"\bREDG?\.E\.ADD\.F32|\bRED\.E\.ADD\.F32"), ("DFMA", r"\bDFMA"), ("HMMA", r"\bHMMA")]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|void |eem::", "", name)
        cur = re.split(r"[<(]", name)[0]
        per.setdefault(cur, [collections.Counter(), 0])
        per[cur][1] += 1
        per[cur].append(collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    for col, pat in COLS:
        if re.search(pat, line):
            per[cur][-1][col] += 1
rows = []
total = collections.Counter()
n_kernels = 0
for k, v in per.items():
    inst = v[2:]
    n_kernels += len(inst)
    best = collections.Counter()
    for c in inst:
        total.update(c)
        for col, _ in COLS:
            best[col] = max(best[col], c[col])
    rows.append((k, len(inst), best))
want = sys.argv[1:] or [k for k, n, b in rows if any(b[c] for c, _ in COLS)]
print("| kernel (instances) | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|" + "---:|" * len(COLS))
for k, n, b in rows:
    if k in want:
        print(f"| `{k}` ({n}) | " + " | ".join(str(b[c]) for c, _ in COLS) + " |")
print(f"| **whole library** ({n_kernels} kernels, sums) | " + " | ".join(str(total[c]) for c, _ in COLS) + " |")
