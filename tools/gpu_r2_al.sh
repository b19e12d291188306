#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2al_q.log
run() { echo "$1" >> gpurun_out/r2al_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel %s kernel_ms %.4f ms_per_step %.3f' % (j['value'], j['e2e']['value'], j['roofline']['kernel'], j['roofline']['kernel_ms'], j['ms_per_step']))
" >> gpurun_out/r2al_q.log; }
EXTRA="--steps 20 --envs 65536"
run "default (8 blocks, 128 regs)" PD_X=1
for m in 6 5 4; do run "minblocks $m" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb$m.so; done
EXTRA="--steps 20 --envs 32768"
run "default 32768" PD_X=1
for m in 6 4; do run "minblocks $m 32768" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb$m.so; done
cat gpurun_out/r2al_q.log
