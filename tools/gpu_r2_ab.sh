#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2ab_q.log
for a in "--steps 30" "--steps 12" "--steps 30 --envs 1024" "--steps 30 --envs 8192" "--steps 20 --car ks_mazda_rx7_tuned"; do timeout 300 python bench.py $a --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step']))
" >> gpurun_out/r2ab_q.log; done
cat gpurun_out/r2ab_q.log
bash tools/gpu_tests.sh r2ab -k "collision or autoreset or single_tick_parity_identical_states or shard or ragged or zero_copy or host_equals"
