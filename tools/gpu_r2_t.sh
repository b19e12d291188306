#!/bin/bash
# A/B: the library built with -fmad=true (FMA contraction everywhere) against the shipped -fmad=false build
mkdir -p gpurun_out
tools/quick_bench.sh 4096 > gpurun_out/r2t_q.log 2>&1
PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_fmad.so tools/quick_bench.sh 4096 >> gpurun_out/r2t_q.log 2>&1
tools/quick_bench.sh 65536 >> gpurun_out/r2t_q.log 2>&1
PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_fmad.so tools/quick_bench.sh 65536 >> gpurun_out/r2t_q.log 2>&1
cat gpurun_out/r2t_q.log
PD_B200_LIB=$PWD/projectd_core_b200/libpd_b200_fmad.so bash tools/gpu_tests.sh r2t
