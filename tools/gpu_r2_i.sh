#!/bin/bash
mkdir -p gpurun_out
tools/quick_bench.sh 65536 > gpurun_out/r2i_q.log 2>&1
PD_SEG_CELL=12 tools/quick_bench.sh 65536 >> gpurun_out/r2i_q.log 2>&1
PD_SEG_CELL=16 tools/quick_bench.sh 65536 >> gpurun_out/r2i_q.log 2>&1
PD_SEG_CELL=16 tools/quick_bench.sh 4096 >> gpurun_out/r2i_q.log 2>&1
PD_SEG_CELL=12 tools/quick_bench.sh 4096 >> gpurun_out/r2i_q.log 2>&1
PD_SEG_CELL=5 tools/quick_bench.sh 4096 >> gpurun_out/r2i_q.log 2>&1
PD_COLL_WARP=0 tools/quick_bench.sh 4096 >> gpurun_out/r2i_q.log 2>&1
cat gpurun_out/r2i_q.log
