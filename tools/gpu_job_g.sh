#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/g_tests.log
timeout 300 python tools/phase_tail.py 4096 2000 > gpurun_out/g_phase.log 2>&1
for e in 1024 4096 16384 65536; do tools/quick_bench.sh $e; done > gpurun_out/g_quick.log 2>&1
cat gpurun_out/g_tests.log; grep -E "tick|arb|probes|factor|sum of" gpurun_out/g_phase.log; cat gpurun_out/g_quick.log
