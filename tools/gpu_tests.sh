#!/bin/bash
# full GPU test tier with a readable log: usage tools/gpu_tests.sh TAG [pytest args]
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -q "$@" > gpurun_out/${tag}_tests_full.log 2>&1
(head -c 1500 gpurun_out/${tag}_tests_full.log; echo; grep -E "passed|failed|Fatal|Error|assert|^E  " gpurun_out/${tag}_tests_full.log | head -40) > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
