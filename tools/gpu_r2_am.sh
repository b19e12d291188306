#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2am_q.log
run() { echo "$1" >> gpurun_out/r2am_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel %s kernel_ms %.4f ms_per_step %.3f' % (j['value'], j['e2e']['value'], j['roofline']['kernel'], j['roofline']['kernel_ms'], j['ms_per_step']))
" >> gpurun_out/r2am_q.log; }
EXTRA="--steps 20 --envs 65536"
run "default (8 blocks, 128 regs)" PD_X=1
run "minblocks 7" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb7.so
run "default again" PD_X=1
run "minblocks 7 again" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb7.so
EXTRA="--steps 20 --envs 24576"
run "default 24576" PD_X=1
run "minblocks 4 24576" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb4.so
run "quad8 24576" PD_QUAD_MAX_ENVS=65536
EXTRA="--steps 20 --envs 40960"
run "default 40960" PD_X=1
run "minblocks 4 40960" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb4.so
run "minblocks 7 40960" PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_mb7.so
cat gpurun_out/r2am_q.log
