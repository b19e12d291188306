#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh lbvh -x -k "device_bvh or synthetic or raycast or fat_points"
PD_DEVICE_BVH=1 timeout 600 python -m pytest tests -m gpu -q -k "raycast_bit_exact or spline_cache or compute_fat_points" 2>&1 | tail -3
(timeout 300 python bench.py --synthetic-tris 1000000 --steps 12 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/lbvh_benchsynth.json
python - <<'PY'
import json
j=json.loads(open('gpurun_out/lbvh_benchsynth.json').read()); print('synthetic 1M (device-built BVH by default): value %.4g e2e %.4g' % (j['value'], j['e2e']['value']))
PY
