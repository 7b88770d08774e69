"""Per-kernel SASS counts of libstrainer_b200.so (cuobjdump -sass | c++filt): the evidence that the hot kernels are tcgen05 / TMA /
TMEM code.   python tools/sass_counts.py > profiles/<round>_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "strainer-gan_b200", "libstrainer_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", ".2CTA", "UTMALDG", "UBLKCP", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "STG.E.ENL2.256", "total"]
rows, cur, it = [], None, iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        rows.append((next(it), cur))
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        continue
    cur["total"] += 1
    for c in cols[:-1]:
        if c == ".2CTA":
            cur[c] += ("UTCHMMA" in line and ".2CTA" in line)
        elif c in line:
            cur[c] += 1
print("per-kernel SASS counts of libstrainer_b200.so (cuobjdump -sass): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG =\n"
      "TMA tensor load / store, UBLKCP = cp.async.bulk (1-D bulk copy), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,\n"
      "STG.256 = 256-bit global stores; kernels without any of them are CUDA-core kernels\n")
print(f"{'kernel':92s}" + "".join(f"{c.replace('STG.E.ENL2.256', 'STG.256'):>9s}" for c in cols))
for name, c in sorted(rows, key=lambda r: (-r[1]["UTCHMMA"], r[0])):
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sg::", "")
    print(f"{short[:91]:92s}" + "".join(f"{c[k]:9d}" for k in cols))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"{'all kernels (' + str(len(rows)) + ')':92s}" + "".join(f"{tot[k]:9d}" for k in cols))
