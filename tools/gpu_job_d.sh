#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/d_tests.log
timeout 300 python tools/warp_tail.py 4096 2000 > gpurun_out/d_tail.log 2>&1
for e in 1024 4096 8192 16384 65536; do tools/quick_bench.sh $e; done > gpurun_out/d_quick.log 2>&1
cat gpurun_out/d_tests.log gpurun_out/d_tail.log gpurun_out/d_quick.log
