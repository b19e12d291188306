#!/bin/bash
mkdir -p gpurun_out
PD_COLL_WARP=0 PD_COLLIDE_V1=1 timeout 300 python tools/collide_tail.py 2>&1 | tail -6 > gpurun_out/r2af_collide.log
PD_COLL_NO_CLIP=1 PD_COLL_WARP=0 PD_COLLIDE_V1=1 timeout 300 python tools/collide_tail.py 2>&1 | tail -6 >> gpurun_out/r2af_collide.log
cat gpurun_out/r2af_collide.log
: > gpurun_out/r2af_q.log
run() { echo "$1" >> gpurun_out/r2af_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f launches %d eps %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step'], j['gpu_launches'], j['episode_stats']['episodes']))
" >> gpurun_out/r2af_q.log; }
for e in "--steps 30" "--steps 30 --envs 65536" "--steps 30 --envs 1024"; do EXTRA="$e"; run "clipped headers $e" PD_X=1; run "unclipped $e" PD_COLL_NO_CLIP=1; done
cat gpurun_out/r2af_q.log
bash tools/gpu_tests.sh r2af -k "collision or autoreset or shard or ragged"
