#!/bin/bash
bash tools/gpu_tests.sh r2j "$@"
(timeout 600 python tools/sac_rollout.py 1024 1500 1 2>&1 | tail -8) > gpurun_out/r2j_sac.log
(timeout 300 python tools/rollout_1024.py 1024 2>&1 | tail -4) > gpurun_out/r2j_rollout.log
cat gpurun_out/r2j_sac.log gpurun_out/r2j_rollout.log
