"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv, --profile-from-start off around ONE step):
groups by kernel (template arguments kept for the tensor-core kernels), prints launches / total us / share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out=None, title=""):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    agg = OrderedDict()
    total = 0.0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        k = r["Kernel Name"]
        k = re.sub(r"\(.*$", "", k).replace("bpb::", "").replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    txt = [f"# {title}", "", f"total {total / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches "
           "(serialised, cold-cache ncu timings: shares are meaningful, absolute times are not)", "",
           "| kernel | launches | total us | share | us / launch |", "|---|---|---|---|---|"]
    for k, (n, us) in rows:
        txt.append(f"| {k} | {n} | {us:.0f} | {100 * us / total:.1f} % | {us / n:.1f} |")
    s = "\n".join(txt) + "\n"
    if out:
        open(out, "w").write(s)
    print(s)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "")
