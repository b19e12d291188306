#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, bench (both arms), plus the secondary configs
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/z_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/z_smoke.log
(timeout 300 python bench.py --impl reference 2>&1 | tail -1) > gpurun_out/z_ref.json
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/z_bench.json
(timeout 300 python bench.py --envs 65536 --steps 10 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/z_bench65536.json
(timeout 300 python bench.py --envs 1024 --steps 12 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/z_bench1024.json
(timeout 300 python bench.py --synthetic-tris 1000000 --steps 12 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/z_benchsynth.json
cat gpurun_out/z_tests.log gpurun_out/z_smoke.log
for f in z_ref z_bench z_bench65536 z_bench1024 z_benchsynth; do python - "$f" <<'PY'
import json,sys
j=json.loads(open('gpurun_out/%s.json'%sys.argv[1]).read())
print(sys.argv[1], 'value %.4g'%j['value'], 'e2e %.4g'%j['e2e']['value'], 'kernel_ms', (j.get('roofline') or {}).get('kernel_ms'), 'launches', j.get('gpu_launches'), 'cpu', (j.get('cpu_baseline') or {}).get('value'))
PY
done
