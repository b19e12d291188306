#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1) > gpurun_out/r2ah_bench2.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 2>&1 | tail -1) > gpurun_out/r2ah_ref2.json
python - <<'PY'
import json
for f in ['r2ah_bench2','r2ah_ref2']:
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read()); print(f, 'n_gpus', j['n_gpus'], 'value %.4g'%j['value'], 'e2e %.4g'%j['e2e']['value'], 'launches', j.get('gpu_launches'))
    except Exception as ex: print(f, 'failed', ex, open('gpurun_out/%s.json'%f).read()[-400:])
PY
