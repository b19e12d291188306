#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2ad_q.log
run() { echo "$1" >> gpurun_out/r2ad_q.log; shift; env "$@" timeout 300 python bench.py --warmup 3 --no-cpu-baseline $EXTRA 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['config']['envs_per_gpu'], 'steps', j['steps'], 'value %.4g e2e %.4g kernel_ms %.4f ms_per_step %.3f launches %d' % (j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step'], j['gpu_launches']))
" >> gpurun_out/r2ad_q.log; }
SEQ=PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_seq.so
for e in "--steps 12" "--steps 30" "--steps 30 --envs 1024" "--steps 30 --envs 8192" "--steps 30 --envs 2048"; do EXTRA="$e"; run "side-by-side $e" PD_X=1; run "sequential $e" $SEQ; done
cat gpurun_out/r2ad_q.log
