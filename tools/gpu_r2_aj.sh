#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2aj_q.log
run() { echo "$1" >> gpurun_out/r2aj_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f launches %d eps %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step'], j['gpu_launches'], j['episode_stats']['episodes']))
" >> gpurun_out/r2aj_q.log; }
for e in "--steps 30 --envs 65536" "--steps 30 --envs 16384"; do EXTRA="$e"; run "k_collide2 $e" PD_X=1; run "floor + walls kernels $e" PD_COLLIDE_SPLIT=1; done
cat gpurun_out/r2aj_q.log
PD_COLLIDE_SPLIT=1 PD_QUAD_MAX_ENVS=0 bash tools/gpu_tests.sh r2aj -k "collision_flag or autoreset or shard"
