#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2a_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/r2a_smoke.log
(timeout 400 python bench.py --steps 12 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2a_bench.json
(timeout 300 python bench.py --envs 65536 --steps 10 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2a_bench65536.json
cat gpurun_out/r2a_tests.log gpurun_out/r2a_smoke.log; tail -c 600 gpurun_out/r2a_bench.json; tail -c 600 gpurun_out/r2a_bench65536.json
