#!/bin/bash
# usage: tools/quick_bench.sh ENVS [bench args]  -> one-line summary of bench.py
e=$1; shift
timeout 300 python bench.py --envs $e --steps 12 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        j=json.loads(l); print(j['config']['envs_per_gpu'], 'value %.4g e2e %.4g kernel %s ms %.4f launches %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel'], j['roofline']['kernel_ms'], j['gpu_launches']))
    except Exception as ex: print(l[:300])
"
