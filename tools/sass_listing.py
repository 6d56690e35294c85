"""profiles/sass_r2.txt: per-kernel counts of the tcgen05 / TMA / TMEM / vector-reduction SASS mnemonics in the shipped
library (cuobjdump -sass; no GPU needed).  Usage: python tools/sass_listing.py > profiles/sass_r2.txt"""
import collections
import re
import subprocess
import sys

LIB = "bpbreid_b200/libbpbreid_sm100.so"
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTCQMMA": r"\bUTCQMMA", "UTMALDG": r"\bUTMALDG", "LDTM": r"\bLDTM", "UTCBAR": r"\bUTCBAR",
       "SYNCS": r"\bSYNCS", "ELECT": r"\bELECT", "REDG.F32x4": r"\bREDG\.E\.ADD\.F32\.FTZ\.RN\.STRONG\.GPU.*\.128|REDG.*\.128",
       "ATOMG/RED": r"\b(ATOMG|REDG|RED)\b"}
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, counts, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[name] = collections.Counter()
        order.append(name)
        continue
    if name:
        for k, p in PAT.items():
            if re.search(p, line):
                counts[name][k] += 1
print("# SASS evidence, round 2 (cuobjdump -sass bpbreid_b200/libbpbreid_sm100.so; counts of the tcgen05 / TMA / TMEM /")
print("# reduction mnemonics per kernel -- regenerate: python tools/sass_listing.py > profiles/sass_r2.txt)")
print()
for n in order:
    c = counts[n]
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print(n)
        print("    " + "  ".join(f"{k}={v}" for k, v in c.items() if v))
print()
print("# kernels without tensor-core / TMA instructions (SIMT: elementwise, reductions, loss, ranking, re-ranking, input pipeline)")
def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("bpb::", "")
    return re.sub(r"<.*", "", re.sub(r"\(.*", "", n))


print("# " + ", ".join(sorted({short(n) for n in order if not (counts[n]["UTCHMMA"] or counts[n]["UTMALDG"] or counts[n]["LDTM"])})))
