#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh
bash scripts/gpu_mel_exp.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pcm.csv python scripts/profile_target.py pcm 1 > gpurun_out/ncu_pcm_list.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_pcm.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; hdr=rows[hi]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=v: continue
    name=r[k].split("(")[0][:60]; a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(r[v])
for n,(c,t) in agg.items(): print(f"  {n:62s} n={c:4d} total={t/1e6:9.3f} ms avg={t/c/1e3:9.2f} us")
PY
