#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -q -x > gpurun_out/r2f_tests_full.log 2>&1
(head -30 gpurun_out/r2f_tests_full.log; grep -E "passed|failed|Fatal|File \"/root|tests/test" gpurun_out/r2f_tests_full.log | head -30) > gpurun_out/r2f_tests.log
tools/quick_bench.sh 4096 > gpurun_out/r2f_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2f_q.log 2>&1
PD_QUAD_CPW=8 tools/quick_bench.sh 8192 >> gpurun_out/r2f_q.log 2>&1
(timeout 300 python tools/phase_tail.py 4096 2000 2>&1 | tail -22) > gpurun_out/r2f_phase.log
cat gpurun_out/r2f_tests.log gpurun_out/r2f_q.log gpurun_out/r2f_phase.log
