#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/f_tests.log
timeout 300 python tools/phase_tail.py 4096 2000 > gpurun_out/f_phase.log 2>&1
for e in 1024 4096 8192 16384 65536; do tools/quick_bench.sh $e; done > gpurun_out/f_quick.log 2>&1
cat gpurun_out/f_tests.log; tail -22 gpurun_out/f_phase.log; cat gpurun_out/f_quick.log
