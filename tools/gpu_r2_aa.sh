#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2050 -c 200 --csv --log-file gpurun_out/r2aa_l1.csv python tools/prof_env.py 4096 2300 > gpurun_out/r2aa_1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 200 --csv --log-file gpurun_out/r2aa_l2.csv python tools/prof_env.py 4096 3150 > gpurun_out/r2aa_2.log 2>&1
python - <<'PY'
import csv
for f in ['r2aa_l1','r2aa_l2']:
    rows=[r for r in csv.reader(open('gpurun_out/%s.csv'%f)) if len(r)>10]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
    v=[float(r[iv].replace(',',''))/1000 for r in rows[1:] if 'k_tick_quad' in r[ik]]
    ev=v[0::2]; od=v[1::2]
    print(f, 'n', len(v), 'mean %.1f' % (sum(v)/len(v)), 'alt-A mean %.1f max %.1f' % (sum(ev)/len(ev), max(ev)), 'alt-B mean %.1f max %.1f' % (sum(od)/len(od), max(od)))
PY
