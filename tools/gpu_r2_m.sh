#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 120 --csv --log-file gpurun_out/r02_launches_65536.csv python tools/prof_env.py 65536 1400 > gpurun_out/r2m_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 120 --csv --log-file gpurun_out/r02_launches_4096.csv python tools/prof_env.py 4096 2400 > gpurun_out/r2m_ncu2.log 2>&1
python - <<'PY'
import csv, collections
for f in ['r02_launches_65536','r02_launches_4096']:
    rows=[r for r in csv.reader(open('gpurun_out/%s.csv'%f)) if len(r)>10]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
    d=collections.defaultdict(list)
    for r in rows[1:]:
        try: d[r[ik].split('(')[0]].append(float(r[iv].replace(',','')))
        except Exception: pass
    print(f, {k:(len(v), round(sum(v)/len(v)/1000,1), round(min(v)/1000,1), round(max(v)/1000,1)) for k,v in d.items()})
PY
