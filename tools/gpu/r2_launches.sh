#!/bin/bash
# ncu launch list of the bench command (per-launch device times; compare shares, not absolutes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0 --sustain-s 0"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMDP > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
ix={k:i for i,k in enumerate(rows[h])}
agg=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)<len(rows[h]): continue
    name=r[ix["Kernel Name"]][:90]; v=float(r[ix["Metric Value"]].replace(",","")); u=r[ix["Metric Unit"]]
    v = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else v)
    agg.setdefault(name, []).append(v)
for k,v in agg.items(): print(f"{len(v):4d} x {sum(v)/len(v):9.3f} ms  {k}")
PY
