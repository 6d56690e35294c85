"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --no-graph`:
takes the LAST forward+loss step (from the last stem_conv_kernel launch on), groups by kernel, prints shares."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    last = max((i for i, (k, _) in enumerate(rows) if "stem_conv" in k), default=0)
    step = rows[last:] if last > 0 else rows
    # cut the probe / e2e launches that follow the step: stop at the next stem conv (none) or at weighted_sum
    for i, (k, _) in enumerate(step):
        if "weighted_sum" in k:
            step = step[:i + 1]
            break
    agg = OrderedDict()
    for k, us in step:
        k = re.sub(r"\(.*", "", k)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    lines = [f"launches in the step: {len(step)}, serialized device time {total / 1e3:.3f} ms (cold-cache, under ncu)", "",
             "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f} % |")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
