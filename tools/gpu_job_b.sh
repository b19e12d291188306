#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/b_tests.log
for c in 8 4 2; do echo "CPW=$c"; PD_QUAD_CPW=$c tools/quick_bench.sh 4096; done > gpurun_out/b_quick.log 2>&1
for c in 4 2; do echo "CPW=$c"; PD_QUAD_CPW=$c tools/quick_bench.sh 1024; done >> gpurun_out/b_quick.log 2>&1
for e in 8192 16384 65536; do tools/quick_bench.sh $e; done >> gpurun_out/b_quick.log 2>&1
echo "quad@16384"; PD_QUAD_MAX_ENVS=20000 tools/quick_bench.sh 16384 >> gpurun_out/b_quick.log 2>&1
cat gpurun_out/b_tests.log gpurun_out/b_quick.log
