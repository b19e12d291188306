#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2ac_q.log
run() { echo "$1" >> gpurun_out/r2ac_q.log; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f launches %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step'], j['gpu_launches']))
" >> gpurun_out/r2ac_q.log; }
EXTRA=""; run "coll warp (default)" PD_X=1
run "k_collide2 lpc4 ahead of the tick" PD_COLL_WARP=0
run "k_collide2 lpc8 ahead" PD_COLL_WARP=0 PD_COLLIDE_LPC=8
run "k_collide v1 ahead" PD_COLL_WARP=0 PD_COLLIDE_V1=1
EXTRA="--envs 1024"; run "1024 default" PD_X=1; run "1024 k_collide2 ahead" PD_COLL_WARP=0
EXTRA="--envs 8192"; run "8192 default" PD_X=1; run "8192 k_collide2 ahead" PD_COLL_WARP=0
cat gpurun_out/r2ac_q.log
