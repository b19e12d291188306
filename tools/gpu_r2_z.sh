#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2z_q.log
for s in 30 30 12 60; do timeout 300 python bench.py --steps $s --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step']))
" >> gpurun_out/r2z_q.log; done
cat gpurun_out/r2z_q.log
