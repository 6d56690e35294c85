#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; tail -12 gpurun_out/probe_conv.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/probe_eval.py > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python - <<'PY'
import csv
rows=[]
for ln in open('gpurun_out/launches_eval.csv'):
    if ln.startswith('"'): rows.append(ln)
agg={}
for r in csv.DictReader(rows):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    k=r["Kernel Name"].split("(")[0]; v=float(r["Metric Value"].replace(",","")); u=r.get("Metric Unit","ns")
    v*= {"ns":1e-3,"us":1.0,"ms":1e3}.get(u,1e-3)
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,us) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{k[:70]:70s} n={n:3d} total={us:10.1f} us avg={us/n:9.1f} us")
PY
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench8.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench8.log | cut -c1-200
