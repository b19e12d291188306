#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2ak_q.log
run() { echo "$1" >> gpurun_out/r2ak_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel %s kernel_ms %.4f ms_per_step %.3f launches %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel'], j['roofline']['kernel_ms'], j['ms_per_step'], j['gpu_launches']))
" >> gpurun_out/r2ak_q.log; }
for n in 12288 16384 24576 32768; do EXTRA="--steps 20 --envs $n"; run "thread per car $n" PD_X=1; run "quad8 $n" PD_QUAD_MAX_ENVS=65536; done
cat gpurun_out/r2ak_q.log
