#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1) > gpurun_out/r02f_bench_${N}gpu.json
python - $N <<'PY'
import json, sys
n=sys.argv[1]
try:
    j=json.loads(open('gpurun_out/r02f_bench_%sgpu.json' % n).read()); print('n_gpus', j['n_gpus'], 'value %.4g'%j['value'], 'e2e %.4g'%j['e2e']['value'], 'ms_per_step %.3f' % j['ms_per_step'], 'launches', j.get('gpu_launches'))
except Exception as ex: print('failed', ex, open('gpurun_out/r02f_bench_%sgpu.json' % n).read()[-600:])
PY
