"""SASS census of libekf_b200.so: per kernel the counts of the instructions that show which hardware path it uses --
DMMA (FP64 tensor pipe), LDGSTS (cp.async), UBLKCP (TMA bulk copy), UTMALDG / UTMASTG (TMA tensor copy), SYNCS (mbarrier),
DFMA, LDS / STS, BAR.  cuobjdump works without a GPU.   python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openekfmonoslam_b200", "lib", "libekf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = collections.OrderedDict([("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"), ("LDGSTS", r"\bLDGSTS"), ("UBLKCP", r"\bUBLKCP"),
                               ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("LDS", r"\bLDS"),
                               ("STS", r"\bSTS"), ("BAR", r"\bBAR\."), ("ACQBULK/ELECT", r"\bELECT|ACQBULK"), ("LDG", r"\bLDG"), ("STG", r"\bSTG")])
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            counts[cur][k] += 1
print(f"SASS census of {os.path.relpath(so, ROOT)} (sm_100a), instruction counts per kernel")
print(f"{'kernel':58s}" + "".join(f"{k:>9s}" for k in PAT))
for name, c in sorted(counts.items()):
    print(f"{name[:57]:58s}" + "".join(f"{c.get(k, 0):9d}" for k in PAT))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':58s}" + "".join(f"{tot.get(k, 0):9d}" for k in PAT))
