#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench14.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench14.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md > /dev/null 2>&1; head -30 gpurun_out/launches_r1.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/probe_eval.py > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python - <<'PY'
import csv
rows=[ln for ln in open('gpurun_out/launches_eval.csv') if ln.startswith('"')]
agg={}
for r in csv.DictReader(rows):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    k=r["Kernel Name"].split("(")[0]; v=float(r["Metric Value"].replace(",","")); u=r.get("Metric Unit","ns")
    v*= {"ns":1e-3,"us":1.0,"ms":1e3}.get(u,1e-3)
    a=agg.setdefault(k,[0,0.0,1e30]); a[0]+=1; a[1]+=v; a[2]=min(a[2],v)
for k,(n,us,mn) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{k[:70]:70s} n={n:3d} total={us:10.1f} us avg={us/n:9.1f} us min={mn:9.1f}")
PY
